"""One turbo-decoder launch for ncu (2048 transport blocks of 1000 bits)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import torch
from config import LTEConfig
from lte_b200 import LinkEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
e2 = LinkEngine.from_config(LTEConfig(5.0, 15.0, 'QPSK'))
plan = e2.coding_plan(1000)
llr = torch.randn((B, plan.sumE), dtype=torch.float32, device='cuda') * 3
e2.tb_decode(llr, plan, want_bits=False)
torch.cuda.synchronize()
