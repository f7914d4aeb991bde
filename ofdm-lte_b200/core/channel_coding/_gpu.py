"""Shared plumbing of the channel-coding stage functions: torch tensors in HBM, the stage-level C entry points
of csrc/coding.cu, no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from lte_b200 import _native as nat
from .. import _backend as be


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream(be.device()).cuda_stream)


def bits_dev(bits):
    return be.as_bits_tensor(bits)                       # uint8 [1, n] on the GPU


def i32(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int32))).to(be.device())


def gather_u8(src, table, n_out):
    out = torch.empty((src.shape[0], n_out), dtype=torch.uint8, device=src.device)
    nat.check(nat.lib.lte_gather_u8(ptr(src), src.shape[1], ptr(table), n_out, ptr(out), src.shape[0], stream()),
              'lte_gather_u8')
    return out


def gather_f32(src, table, n_out):
    out = torch.empty((src.shape[0], n_out), dtype=torch.float32, device=src.device)
    nat.check(nat.lib.lte_gather_f32(ptr(src), src.shape[1], ptr(table), n_out, ptr(out), src.shape[0], stream()),
              'lte_gather_f32')
    return out


def gather_any(data, table_np):
    """data[table] on the GPU for bit arrays (uint8 kernel) or real / complex values (float32 kernel)."""
    a = np.asarray(data)
    t = i32(table_np)
    if np.iscomplexobj(a):
        re = gather_f32(torch.from_numpy(a.real.astype(np.float32)[None]).to(be.device()), t, len(table_np))
        im = gather_f32(torch.from_numpy(a.imag.astype(np.float32)[None]).to(be.device()), t, len(table_np))
        return (re + 1j * im).cpu().numpy()[0].astype(a.dtype)
    if a.dtype.kind == 'f':
        return gather_f32(torch.from_numpy(a.astype(np.float32)[None]).to(be.device()), t, len(table_np)) \
            .cpu().numpy()[0].astype(a.dtype)
    if a.size and (a.min() < 0 or a.max() > 255):
        raise ValueError("integer data outside 0..255 cannot go through the bit kernels")
    return gather_u8(torch.from_numpy(a.astype(np.uint8)[None]).to(be.device()), t, len(table_np)) \
        .cpu().numpy()[0].astype(a.dtype)


def single_block_tables(K, with_pi=True):
    """blk table and QPP permutation of one code block of K bits (no filler, no CRC)."""
    from lte_b200.qpp_table import QPP
    blk = np.zeros((1, nat.LTE_BLK_COLS), dtype=np.int32)
    blk[0] = (K, 0, K, 0, 0, 0, 0, 0)
    if with_pi:
        f1, f2 = QPP[K]
        i = np.arange(K, dtype=np.int64)
        pi = ((f1 * i + f2 * i * i) % K).astype(np.int32)
    else:
        pi = np.arange(max(K, 1), dtype=np.int32)
    return i32(blk), i32(pi)
