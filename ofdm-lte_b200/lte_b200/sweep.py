"""Monte-Carlo BER sweeps sharded over GPUs.

Every (trial, SNR point) is an independent link realisation ("stream").  Stream ids are global:
    stream_id = trial * n_snr + snr_index
and every random draw (bits, Jakes phases, noise) is keyed by (seed, stream_id[, antenna]), so the
summed error counts do not depend on the batch size or on how many GPUs share the trial range.
Ranks own contiguous trial ranges; the only collective is one all-reduce of the int64
[2, n_snr] counters (errors, bits) at the end -- the replacement for the reference's per-antenna
thread pool (core/parallel_processing.py:23-223, core/ofdm_core.py:1615-1630).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous partition of range(n) over `world` ranks; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counts(counts):
    """Sum an int64 counter tensor over all ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def run_sweep(count_batch, n_snr, n_trials, bits_per_stream, batch_trials=256, rank=0, world=1, device='cpu'):
    """Generic driver.  count_batch(trial_lo, n) -> int64 tensor [n * n_snr] of bit errors for
    trials trial_lo .. trial_lo + n - 1 (stream-major: trial outer, SNR inner).
    Returns dict(errors[n_snr], bits[n_snr], ber[n_snr]) reduced over all ranks."""
    lo, hi = shard_range(n_trials, rank, world)
    counts = torch.zeros((2, n_snr), dtype=torch.int64, device=device)
    t = lo
    while t < hi:
        n = min(batch_trials, hi - t)
        err = count_batch(t, n)
        counts[0] += err.view(n, n_snr).sum(dim=0).to(counts.device)
        counts[1] += n * int(bits_per_stream)
        t += n
    reduce_counts(counts)
    errors, bits = counts[0].cpu(), counts[1].cpu()
    return {'errors': errors, 'bits': bits, 'ber': errors.double() / bits.clamp(min=1).double()}


def simo_sweep(engine, chan, snr_db, n_trials, num_rx, symbols_per_stream=14, seed=0, batch_trials=256,
               rank=0, world=1, noise_domain=1, fused=False):
    """BER of the SIMO-MRC chain at every SNR point, `n_trials` independent streams per point.
    noise_domain / fused: see LinkEngine.simo_ber."""
    n_snr = len(snr_db)
    S, R = symbols_per_stream, num_rx
    snr_lin = torch.tensor([10 ** (s / 10) for s in snr_db], dtype=torch.float32, device=engine.device)
    state = {}

    def count_batch(trial_lo, n):
        B = n * n_snr
        if state.get('B') != B:
            state['B'] = B
            state['ws'] = engine.workspace(B, S, R, fading=chan.num_taps > 0, fused=fused)
            state['snr_rows'] = snr_lin.repeat(n).repeat_interleave(R).contiguous()
        return engine.simo_ber(state['ws'], chan, state['snr_rows'], seed, stream_id0=trial_lo * n_snr,
                               noise_domain=noise_domain, fused=fused).clone()

    return run_sweep(count_batch, n_snr, n_trials, S * engine.Nd * engine.bps, batch_trials, rank, world,
                     engine.device)


def beamforming_sweep(engine, codebook, snr_db, n_trials, num_tx, num_rx, mode='MRT', symbols_per_stream=14, seed=0,
                      batch_trials=1024, rank=0, world=1):
    """BER of the rank-1 beamforming link (reference OFDMSimulator.simulate_beamforming,
    core/ofdm_core.py:2260-2477) at every SNR point: each stream draws its own flat R x T channel,
    selects its precoder (mode 'MRT' = update_mode 'adaptive', 'CODEBOOK' = 'static'), and runs
    S OFDM symbols through x = W s, y = H x + n, MRC, slicer, count.  Also returns the mean array
    gain in dB and the PMI histogram, reduced like the error counters."""
    n_snr = len(snr_db)
    S, R, T = symbols_per_stream, num_rx, num_tx
    nstd = torch.tensor([(10 ** (-s / 10) / 2) ** 0.5 for s in snr_db], dtype=torch.float32, device=engine.device)
    ncb = len(codebook)
    extra = torch.zeros(2 + ncb, dtype=torch.float64, device=engine.device)     # sum gain_db, streams, PMI histogram

    def count_batch(trial_lo, n):
        B = n * n_snr
        sid0 = trial_lo * n_snr
        idx = engine.random_indices(B, S, seed, sid0)
        h = engine.random_channel(B, R, T, seed, sid0)
        W, heff, pmi, gain = engine.bf_weights(h, codebook, mode=mode)
        err, _ = engine.bf_link(idx, h, W, heff, nstd.repeat(n).contiguous(), S, seed=seed, row_id0=sid0 * R)
        extra[0] += gain.double().sum()
        extra[1] += B
        extra[2:] += torch.bincount(pmi.long(), minlength=ncb).double()
        return err

    out = run_sweep(count_batch, n_snr, n_trials, S * engine.Nd * engine.bps, batch_trials, rank, world, engine.device)
    reduce_counts(extra)
    out['mean_gain_db'] = float(extra[0] / extra[1].clamp(min=1))
    out['pmi_hist'] = extra[2:].long().cpu()
    return out
