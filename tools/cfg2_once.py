"""One batch of BASELINE config 2 (SISO SC-FDM 10 MHz 16-QAM Pedestrian_A, BER + PAPR histogram) for a launch list."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import torch  # noqa: E402
from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for  # noqa: E402

cfg = LTEConfig(10.0, 15.0, '16-QAM')
eng = LinkEngine.from_config(cfg)
chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
B = 8192
snr = torch.tensor([10 ** (s / 10) for s in range(0, 31, 2)], dtype=torch.float32, device='cuda').repeat(B // 16).contiguous()
hist = torch.zeros(200, dtype=torch.int64, device='cuda')
for _ in range(2):
    eng.siso_ber(chan, snr, 14, 1, sc_fdm=True, papr_hist=hist)
torch.cuda.synchronize()
print('ok')
