"""One batch of BASELINE config 4 (2x2 SFBC) and config 5 (4x4 SM) for a launch-list capture."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import torch  # noqa: E402
from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for, tables  # noqa: E402

dev = torch.device('cuda', 0)
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
snr16 = torch.tensor([10 ** (s / 10) for s in range(0, 31, 2)], dtype=torch.float32, device=dev)
if which in ('4', 'both'):
    cfg = LTEConfig(20.0, 15.0, '16-QAM')
    eng0 = LinkEngine.from_config(cfg, device=dev)
    eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(2, eng0.Np), device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    B = 2048
    rows = snr16.repeat(B // 16).repeat_interleave(2).contiguous()
    for _ in range(2):
        eng.sfbc_ber(chan, rows, 14, 2, 1)
if which in ('5', 'both'):
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng0 = LinkEngine.from_config(cfg, device=dev)
    eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(4, eng0.Np), device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
    from core.codebook_lte import LTECodebook
    W = LTECodebook(4, transmission_mode='TM4', rank=4).get_precoder(0)
    for _ in range(2):
        eng.sm_ber(chan, W, 20.0, 512, 14, 4, 'MMSE', 1)
torch.cuda.synchronize()
print('ok')
