"""BCJR turbo decoder (reference core/channel_coding/turbo_decoder.py:24-446) on `lte_turbo_decode_blocks`:
max-log-MAP (the reference's default, USE_MAX_LOG_MAP = True) or, after set_decoder_mode(False), the exact
Jacobian logarithm max*(a, b) = max + log1p(e^-|a-b|)."""
import numpy as np
import torch

from lte_b200 import _native as nat
from lte_b200.qpp_table import QPP
from . import _gpu as g
from .turbo_encoder import qpp_deinterleave, qpp_interleave  # noqa: F401  (re-exported by the reference module)

USE_MAX_LOG_MAP = True


def set_decoder_mode(use_max_log_map=True):
    global USE_MAX_LOG_MAP
    USE_MAX_LOG_MAP = bool(use_max_log_map)
    print(f"Turbo Decoder mode set to: {'Max-Log-MAP (fast)' if USE_MAX_LOG_MAP else 'True Log-MAP (exact)'}")


def log_sum_exp(a, b):
    if np.isinf(a) and a < 0:
        return b
    if np.isinf(b) and b < 0:
        return a
    return max(a, b) + np.log1p(np.exp(-abs(a - b)))


def max_star(a, b):
    return max(a, b) if USE_MAX_LOG_MAP else log_sum_exp(a, b)


def _run(dl, K, iterations, apriori=None, want_apost=False, with_pi=True):
    blk, pi = g.single_block_tables(K, with_pi)
    dev = g.be.device()
    x = torch.from_numpy(np.ascontiguousarray(dl, dtype=np.float32)[None]).to(dev)
    work = torch.empty((4, int(nat.lib.lte_tb_decode_work_floats(K))), dtype=torch.float32, device=dev)
    bits = torch.empty((1, K), dtype=torch.uint8, device=dev)
    ap = torch.from_numpy(np.ascontiguousarray(apriori, dtype=np.float32)[None]).to(dev) if apriori is not None else None
    apost = torch.empty((1, K + 3), dtype=torch.float32, device=dev) if want_apost else None
    nat.check(nat.lib.lte_turbo_decode_blocks(g.ptr(x), g.ptr(blk), 1, K, 3 * K + 12, K, g.ptr(pi), int(iterations),
                                              0 if USE_MAX_LOG_MAP else 1, g.ptr(work), g.ptr(bits), g.ptr(ap), g.ptr(apost), 1, g.stream()),
              'lte_turbo_decode_blocks')
    return bits.cpu().numpy()[0], (apost.cpu().numpy()[0].astype(np.float64) if want_apost else None)


def turbo_decode(llr_encoded, K, num_iterations=5, debug=False):
    """3K + 12 LLRs in encoder order -> K hard decisions after `num_iterations` iterations + the final pass."""
    if K not in QPP:
        raise ValueError(f"Invalid interleaver size K={K}")
    llr = np.asarray(llr_encoded, dtype=np.float64)
    if len(llr) < 3 * K + 12:
        llr = np.concatenate([llr, np.zeros(3 * K + 12 - len(llr))])
    bits, _ = _run(llr[:3 * K + 12], K, num_iterations)
    return bits


class LogMAPDecoder:
    """One BCJR pass over an 8-state RSC trellis that starts and ends in state 0 (reference :100-337)."""

    def __init__(self):
        self.num_states = 8
        self.num_memory = 3

    def decode(self, llr_systematic, llr_parity, llr_apriori=None, return_extrinsic=True):
        Ls = np.asarray(llr_systematic, dtype=np.float64)
        Lp = np.asarray(llr_parity, dtype=np.float64)
        n = len(Ls)
        La = np.zeros(n) if llr_apriori is None else np.asarray(llr_apriori, dtype=np.float64)
        if n < 4:
            raise ValueError("the CUDA BCJR pass needs at least 4 trellis steps")
        if np.any(La[n - 3:] != 0):
            raise ValueError("a-priori LLRs on the last three (termination) steps are not supported")
        K = n - 3
        dl = np.zeros(3 * K + 12)
        dl[0:3 * K:3], dl[1:3 * K:3] = Ls[:K], Lp[:K]
        dl[3 * K:3 * K + 3], dl[3 * K + 3:3 * K + 6] = Ls[K:], Lp[K:]
        _, apost = _run(dl, K, 0, apriori=La[:K], want_apost=True, with_pi=False)
        bits = (apost < 0).astype(np.uint8)
        # float32 inputs were used on the device: form the extrinsic from the same rounded values
        out = apost - La.astype(np.float32) - Ls.astype(np.float32) if return_extrinsic else apost
        return bits, out
