"""Sub-block interleaver and circular-buffer rate matching (reference core/channel_coding/rate_matching.py:25-472)
as index tables applied by `lte_gather_u8` / `lte_gather_f32`.  The reference's own variant is kept: the 32-column
matrix is filled column by column with the nulls at the END, columns are permuted by bit reversal, rows are read
out and nulls dropped."""
import numpy as np
import torch

from lte_b200.coding import _sub_block_order
from . import _gpu as g


def sub_block_interleaver(input_bits, D=32):
    a = np.asarray(input_bits)
    if len(a) == 0:
        return np.array([], dtype=np.uint8)
    return g.gather_any(a, _sub_block_order(len(a)).astype(np.int32)).astype(np.uint8)


def sub_block_deinterleaver(input_bits, original_length, D=32):
    if original_length == 0:
        return np.array([], dtype=np.uint8)
    order = _sub_block_order(original_length)
    inv = np.empty(original_length, dtype=np.int32)
    inv[order] = np.arange(original_length, dtype=np.int32)
    return g.gather_any(np.asarray(input_bits)[:original_length], inv).astype(np.uint8)


def sub_block_deinterleaver_llr(interleaved_data, K_original):
    a = np.asarray(interleaved_data, dtype=np.float64)
    order = _sub_block_order(len(a))
    inv = np.empty(len(a), dtype=np.int32)
    inv[order] = np.arange(len(a), dtype=np.int32)
    return g.gather_any(a, inv)[:K_original]


def _circular_buffer_source(K):
    """cbuf[p] = index into the 3K+12 encoded bits, -1 for the padding of the two parity streams."""
    k = np.arange(K)
    streams = [np.concatenate([3 * k, 3 * K + np.arange(3), 3 * K + 6 + np.arange(3)]),
               np.concatenate([3 * k + 1, 3 * K + 3 + np.arange(3)]),
               np.concatenate([3 * k + 2, 3 * K + 9 + np.arange(3)])]
    buf = np.full((K + 6, 3), -1, dtype=np.int64)
    for j, d in enumerate(streams):
        v = d[_sub_block_order(len(d))]
        buf[:len(v), j] = v
    return buf.reshape(-1)


def rate_match_turbo(encoded_bits, E, K, rv_idx=0):
    enc = np.asarray(encoded_bits)
    if len(enc) != 3 * K + 12:
        raise ValueError(f"Invalid encoded_bits length. Expected {3*K + 12}, got {len(enc)}")
    cb = _circular_buffer_source(K)
    N = len(cb)
    start = [0, N // 4, N // 2, 3 * N // 4][rv_idx % 4]
    table = cb[(start + np.arange(E)) % N].astype(np.int32)
    return g.gather_any(enc.astype(np.uint8), table).astype(np.uint8)


def rate_dematching_turbo(rate_matched_llrs, K, rv_idx=0, debug=False):
    llr = np.asarray(rate_matched_llrs, dtype=np.float64)
    E = len(llr)
    cb = _circular_buffer_source(K)
    N = len(cb)
    start = [0, N // 4, N // 2, 3 * N // 4][rv_idx % 4]
    pos = (start + np.arange(E)) % N
    dev = g.be.device()
    # repeated positions accumulate (reference :323-326); torch's index_add_ on the device
    buf = torch.zeros(N, dtype=torch.float32, device=dev)
    buf.index_add_(0, torch.from_numpy(pos).to(dev), torch.from_numpy(llr.astype(np.float32)).to(dev))
    table = np.full(3 * K + 12, -1, dtype=np.int32)
    ok = cb >= 0
    table[cb[ok]] = np.flatnonzero(ok)
    return g.gather_f32(buf[None], g.i32(table), 3 * K + 12).cpu().numpy()[0].astype(np.float64)
