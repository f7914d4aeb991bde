"""Latency of one coded transport block of an image-sized payload (1.62 Mbit, 270 code blocks) through
simulate_siso_coded's device chain."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import numpy as np, torch
from config import LTEConfig
from lte_b200 import LinkEngine, chan_for
eng = LinkEngine.from_config(LTEConfig(10.0, 15.0, '64-QAM'))
A = 450 * 450 * 3 * 8
bits = torch.from_numpy(np.random.RandomState(0).randint(0, 2, (1, A)).astype(np.uint8)).cuda()
plan = eng.coding_plan(A)
awgn = chan_for('awgn', eng.fs, 'Pedestrian_A', 2.0, 0.0)
snr = torch.full((1,), 22.0, dtype=torch.float32, device='cuda')
for name, fn in (('tb_encode', lambda: eng.tb_encode(bits, plan)), ('chain', lambda: eng.siso_coded_ber(bits, awgn, snr, 1, 0))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print(name, round((time.perf_counter() - t0) * 1e3, 2), 'ms', 'C =', plan.C)
err, crc = out
print('errors', int(err.sum()), 'crc', int(crc[0]))
