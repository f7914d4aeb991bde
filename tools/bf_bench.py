import sys, os, json
sys.path[:0] = ['/root/repo', '/root/repo/ofdm-lte_b200', '/root/repo/tools']
import torch
import extras_bench as xb
from config import LTEConfig
from lte_b200 import LinkEngine
from core.codebook_lte import LTECodebook
eng = LinkEngine.from_config(LTEConfig(10.0, 15.0, '64-QAM'))
for (T, R) in ((4, 2), (8, 1), (2, 2)):
    cb = LTECodebook(T, 'TM6').codebook
    B, S = 8192, 14
    idx = eng.random_indices(B, S, 1, 0); h = eng.random_channel(B, R, T, 1, 0)
    nstd = torch.full((B,), 0.1, dtype=torch.float32, device='cuda')
    W, heff, pmi, gain = eng.bf_weights(h, cb, mode='CODEBOOK')
    print(T, R, round(xb.timed(lambda: eng.bf_link(idx, h, W, heff, nstd, S, seed=1), 5), 4), 'ms')
