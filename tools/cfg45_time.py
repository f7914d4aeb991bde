"""Device time of one batch of BASELINE config 4 (2x2 SFBC) and config 5 (4x4 MMSE SM), fused link vs staged kernels."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import torch  # noqa: E402
from config import LTEConfig  # noqa: E402
from core.codebook_lte import LTECodebook  # noqa: E402
from lte_b200 import LinkEngine, chan_for, tables  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cfg = LTEConfig(20.0, 15.0, '16-QAM')
eng0 = LinkEngine.from_config(cfg)
eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(2, eng0.Np))
chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
snr16 = torch.tensor([10 ** (s / 10) for s in range(0, 31, 2)], dtype=torch.float32, device='cuda')
rows = snr16.repeat(2048 // 16).repeat_interleave(2).contiguous()
for fused in (False, True):
    print('config 4, 2048 subframes, fused', fused, round(timed(lambda: eng.sfbc_ber(chan, rows, 14, 2, 1, fused=fused)), 3), 'ms')
cfg = LTEConfig(20.0, 15.0, '64-QAM')
eng0 = LinkEngine.from_config(cfg)
eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(4, eng0.Np))
chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
W = LTECodebook(4, transmission_mode='TM4', rank=4).get_precoder(0)
for fused in (False, True):
    print('config 5, 512 subframes, fused', fused, round(timed(lambda: eng.sm_ber(chan, W, 20.0, 512, 14, 4, 'MMSE', 1, fused=fused)), 3), 'ms')
