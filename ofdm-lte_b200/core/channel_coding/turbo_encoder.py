"""LTE turbo encoder (reference core/channel_coding/turbo_encoder.py:27-330) on `lte_turbo_encode_blocks`; the
QPP (de)interleavers are `lte_gather_*` with the permutation as a table."""
import numpy as np
import torch

from lte_b200 import _native as nat
from lte_b200.qpp_table import QPP
from . import _gpu as g

QPP_INTERLEAVER_PARAMS = dict(QPP)


def _perm(K):
    if K not in QPP:
        raise ValueError(f"Invalid interleaver size K={K}")
    f1, f2 = QPP[K]
    i = np.arange(K, dtype=np.int64)
    return ((f1 * i + f2 * i * i) % K).astype(np.int32)


def qpp_interleave(data, K):
    return g.gather_any(np.asarray(data)[:K], _perm(K))


def qpp_deinterleave(data, K):
    p = _perm(K)
    inv = np.empty(K, dtype=np.int32)
    inv[p] = np.arange(K, dtype=np.int32)
    return g.gather_any(np.asarray(data)[:K], inv)


def turbo_encode(input_bits):
    """K bits -> 3K + 12: [s_k, p1_k, p2_k] interleaved, then tail sys1, par1, sys2, par2 (3 each)."""
    u = np.asarray(input_bits)
    K = len(u)
    if K not in QPP:
        raise ValueError(f"Invalid code block size K={K}. Must be valid interleaver size.")
    blk, pi = g.single_block_tables(K)
    cb = g.bits_dev(u)
    enc = torch.empty((1, 3 * K + 12), dtype=torch.uint8, device=cb.device)
    nat.check(nat.lib.lte_turbo_encode_blocks(g.ptr(cb), g.ptr(blk), 1, K, 3 * K + 12, g.ptr(pi), g.ptr(enc), 1,
                                              g.stream()), 'lte_turbo_encode_blocks')
    return enc.cpu().numpy()[0]


def rsc_encode(input_bits, trellis_termination=True):
    """One constituent encoder (reference :137-211): (systematic = feedback bits, parity), 3 tail steps.  Runs the
    turbo-encoder kernel on a block padded to a valid size and reads constituent encoder 1's streams."""
    u = np.asarray(input_bits).astype(np.uint8)
    n = len(u)
    from lte_b200.coding import interleaver_size
    K = interleaver_size(max(n, 40))
    # zero padding AFTER the data keeps encoder 1's first n outputs; its state after n steps gives the tail
    enc = turbo_encode(np.concatenate([u, np.zeros(K - n, dtype=np.uint8)]))
    sys_, par = enc[0:3 * n:3].copy(), enc[1:3 * n:3].copy()
    if not trellis_termination:
        return sys_, par
    s0 = int(sys_[-1]) if n >= 1 else 0
    s1 = int(sys_[-2]) if n >= 2 else 0
    s2 = int(sys_[-3]) if n >= 3 else 0
    ts, tp = [], []
    for _ in range(3):                               # three flush steps of a 3-bit register (scalar bookkeeping)
        fb = 0                                       # tail input s1 ^ s2 makes the feedback 0
        ts.append(fb)
        tp.append(fb ^ s0 ^ s2)
        s0, s1, s2 = fb, s0, s1
    return np.concatenate([sys_, np.array(ts, dtype=np.uint8)]), np.concatenate([par, np.array(tp, dtype=np.uint8)])


def turbo_encode_block_list(code_blocks):
    return [turbo_encode(b) for b in code_blocks]
