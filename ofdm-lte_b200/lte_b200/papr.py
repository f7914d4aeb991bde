"""PAPR / CCDF sweep engine.

Replaces the reference's collection loop `OFDMSystem.collect_papr_for_all_modulations`
(core/ofdm_system.py:648-735), which modulates `n_simulations` random bit vectors per
(modulation, OFDM | SC-FDM) pair, computes the per-symbol PAPR of the useful part
(`calculate_papr_without_cp`, :173-229) and hands the values to a CCDF plot.

Here one launch sequence per batch does: Philox symbol indices -> (SC-FDM: constellation map +
M-point unitary DFT) -> grid + IFFT with the PAPR epilogue.  The time-domain stream is never
written; the only outputs are the histogram (and, on request, the per-symbol values).  Symbol
ids are global (`stream_id0`), so the histogram does not depend on batch size or GPU count;
ranks own contiguous ranges and the single collective is one all-reduce of the int64 histogram.
"""
import numpy as np
import torch

from .sweep import reduce_counts, shard_range


def ccdf_from_hist(hist, lo, step):
    """-> (thresholds_db [bins], P(PAPR > threshold) [bins]) at the upper edge of every bin but the
    overflow bin, whose threshold is +inf-like (its own upper edge) with probability 0."""
    h = np.asarray(hist, dtype=np.float64)
    total = h.sum()
    edges = lo + step * np.arange(1, h.size + 1)
    if total == 0:
        return edges, np.zeros_like(edges)
    above = total - np.cumsum(h)
    return edges, above / total


def papr_sweep(engine, n_streams, symbols_per_stream=14, sc_fdm=False, seed=0, batch_streams=1024,
               hist_lo=0.0, hist_step=0.05, hist_bins=400, rank=0, world=1, return_values=False):
    """Histogram of the per-OFDM-symbol PAPR (dB, no CP) of `n_streams` random streams of
    `symbols_per_stream` symbols with the engine's numerology and modulation.

    Returns dict(hist int64 [bins] summed over ranks, count, lo, step, thresholds_db, ccdf,
    mean_db, max_db[, values float32 [local symbols] when return_values])."""
    S = int(symbols_per_stream)
    lo_s, hi_s = shard_range(n_streams, rank, world)
    dev = engine.device
    hist = torch.zeros(hist_bins, dtype=torch.int64, device=dev)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)        # sum of dB values, local count
    mx = torch.full((1,), -float('inf'), dtype=torch.float32, device=dev)
    values = []
    t = lo_s
    while t < hi_s:
        n = min(batch_streams, hi_s - t)
        idx = engine.random_indices(n, S, seed, stream_id0=t)
        if sc_fdm:
            pre = engine.dft_m(engine.qam_map(idx).reshape(n * S, engine.Nd), engine.Nd)
            db, _, _ = engine.modulate_papr(S, symbols=pre.reshape(n, S * engine.Nd), hist=hist, hist_lo=hist_lo,
                                            hist_step=hist_step)
        else:
            db, _, _ = engine.modulate_papr(S, idx=idx, hist=hist, hist_lo=hist_lo, hist_step=hist_step)
        acc[0] += db.double().sum()
        acc[1] += db.numel()
        mx = torch.maximum(mx, db.max().reshape(1))
        if return_values:
            values.append(db.reshape(-1).clone())
        t += n
    reduce_counts(hist)
    reduce_counts(acc)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
    h = hist.cpu().numpy()
    thr, ccdf = ccdf_from_hist(h, hist_lo, hist_step)
    count = int(acc[1].item())
    out = {'hist': h, 'count': count, 'lo': hist_lo, 'step': hist_step, 'thresholds_db': thr, 'ccdf': ccdf,
           'mean_db': float(acc[0].item() / max(count, 1)), 'max_db': float(mx.item())}
    if return_values:
        out['values'] = torch.cat(values) if values else torch.empty(0, dtype=torch.float32, device=dev)
    return out
