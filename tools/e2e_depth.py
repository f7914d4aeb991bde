"""End-to-end (host bits -> errors on host) subframes/s of the headline batch for several pipeline depths.
usage: python tools/e2e_depth.py [B] [batches]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device('cuda', 0)
cfg = LTEConfig(20.0, 15.0, '64-QAM')
eng = LinkEngine.from_config(cfg, device=dev)
chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
S, R = 14, 4
nbits = S * eng.Nd * eng.bps
snr = torch.tensor([10 ** (s / 10) for s in range(0, 31, 2)], dtype=torch.float32, device=dev)
rows = snr.repeat(B // 16).repeat_interleave(R).contiguous()
host_bits = torch.from_numpy(np.random.RandomState(7).randint(0, 256, (B, (nbits + 7) // 8), dtype=np.uint8)).pin_memory()
for depth in (2, 3, 4):
    pipe = eng.stream_host_batches(chan, R, rows, B, S, nbits=nbits, seed=1, noise_domain=3, depth=depth)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        tot = 0
        for err, _ in pipe.run((host_bits, k * B, k) for k in range(nb)):
            tot += int(err[0])
        e1.record()
        torch.cuda.synchronize()
    print(f'depth {depth}: {B * nb / (e0.elapsed_time(e1) * 1e-3) / 1e6:.3f} M subframes/s ({e0.elapsed_time(e1) / nb:.3f} ms/batch)')
