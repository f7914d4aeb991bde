"""Host-side layout of a coded transport block.

Everything here depends only on the transport-block length A, so it is computed once per length and
uploaded as small int32 tables; the per-bit work (CRC shift registers, RSC encoders, gathers, soft
demapping, BCJR recursions) runs in csrc/coding.cu.  Reference: core/channel_coding/segmentation.py
(:66-199 block sizes K+/K-, filler, bit split), rate_matching.py (:27-229 sub-block interleaver in
the reference's own column-fill variant, circular buffer with E = 3K+12 and rv 0) and
turbo_encoder.py (:74-109 QPP interleaver)."""
import numpy as np
import torch

from . import _native as nat
from .qpp_table import QPP, TURBO_K

_Z, _L = 6144, 24
_COL_PERM = np.array([int(f'{c:05b}'[::-1], 2) for c in range(32)])     # bit-reversed column order (rate_matching.py:39-42)


def interleaver_size(n):
    for k in TURBO_K:
        if k >= n:
            return k
    raise ValueError(f"No valid interleaver size found for min_size={n}")


def block_layout(B):
    """[(K, filler, info bits, has_crc24b)] for a transport block of B bits including its CRC-24A."""
    if B <= _Z:
        K = interleaver_size(B)
        return [(K, K - B, B, 0)]
    C = -(-B // (_Z - _L))
    Bp = B + C * _L
    Kp = interleaver_size(-(-Bp // C))
    i = TURBO_K.index(Kp)
    Km = TURBO_K[i - 1] if i > 0 else Kp
    Cm = (C * Kp - Bp) // (Kp - Km) if Kp > Km else 0
    out, left = [], B
    for r in range(C):
        K = Km if r < Cm else Kp
        n = left if r == C - 1 else min(K - _L, left // (C - r))
        left -= n
        out.append((K, K - _L - n, n, 1))
    return out


def _sub_block_order(n):
    """Read-out order of one stream through the sub-block interleaver: element e sits at row e % R,
    column e // R; columns are visited in bit-reversed order, rows outermost, nulls skipped."""
    R = -(-n // 32)
    e = _COL_PERM[None, :] * R + np.arange(R)[:, None]
    e = e.reshape(-1)
    return e[e < n]


def rate_match_source(K):
    """coded[i] = encoded[src[i]] (-1: constant 0) for E = 3K + 12, rv 0."""
    k = np.arange(K)
    streams = [np.concatenate([3 * k, 3 * K + np.arange(3), 3 * K + 6 + np.arange(3)]),      # systematic + both tails
               np.concatenate([3 * k + 1, 3 * K + 3 + np.arange(3)]),                          # parity 1 + tail
               np.concatenate([3 * k + 2, 3 * K + 9 + np.arange(3)])]                          # parity 2 + tail
    width = K + 6
    buf = np.full((width, 3), -1, dtype=np.int64)
    for j, d in enumerate(streams):
        v = d[_sub_block_order(len(d))]
        buf[:len(v), j] = v
    return buf.reshape(-1)[:3 * K + 12]


class CodingPlan:
    """Tables of one transport-block length A on one device."""

    def __init__(self, A, device):
        self.A = int(A)
        if self.A < 1:
            raise ValueError("Bits array cannot be empty")
        self.layout = block_layout(self.A + 24)
        self.C = len(self.layout)
        Ks = [k for k, _, _, _ in self.layout]
        self.Kmax = max(Ks)
        self.sumK = int(sum(Ks))
        self.sumE = int(sum(3 * k + 12 for k in Ks))
        pi_off, pis, off = {}, [], 0
        for K in sorted(set(Ks)):
            f1, f2 = QPP[K]
            i = np.arange(K, dtype=np.int64)
            pis.append(((f1 * i + f2 * i * i) % K).astype(np.int32))
            pi_off[K] = off
            off += K
        blk = np.zeros((self.C, nat.LTE_BLK_COLS), dtype=np.int32)
        rm = np.empty(self.sumE, dtype=np.int32)
        dm = np.full(self.sumE, -1, dtype=np.int32)
        src = cb = enc = 0
        for r, (K, F, n, has_crc) in enumerate(self.layout):
            blk[r] = (K, F, n, src, cb, enc, has_crc, pi_off[K])
            t = rate_match_source(K)
            E = 3 * K + 12
            rm[enc:enc + E] = np.where(t >= 0, t + enc, -1)
            ok = t >= 0
            dm[enc + t[ok]] = enc + np.flatnonzero(ok)
            src, cb, enc = src + n, cb + K, enc + E
        dev = torch.device(device)
        self.blk = torch.from_numpy(blk).to(dev)
        self.rm_table = torch.from_numpy(rm).to(dev)
        self.dm_table = torch.from_numpy(dm).to(dev)
        self.pi_tab = torch.from_numpy(np.concatenate(pis)).to(dev)
        self.rm_host, self.dm_host = rm, dm
        self.work_floats = int(nat.lib.lte_tb_decode_work_floats(self.Kmax))
