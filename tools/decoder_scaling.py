import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/ofdm-lte_b200']
import torch, numpy as np
from config import LTEConfig
from lte_b200 import LinkEngine
e2 = LinkEngine.from_config(LTEConfig(5.0, 15.0, 'QPSK'))
for A in (1000, 6120):
    plan = e2.coding_plan(A)
    for B in (1024, 4096, 16384, 65536):
        if B * plan.C * plan.work_floats * 4 > 20e9: continue
        llr = torch.randn((B, plan.sumE), dtype=torch.float32, device='cuda') * 3
        e2.tb_decode(llr, plan, want_bits=False); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e2.tb_decode(llr, plan, want_bits=False); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(A, B, round(ms, 2), 'ms', round(B * A / ms / 1e3, 1), 'Mbit/s info')
