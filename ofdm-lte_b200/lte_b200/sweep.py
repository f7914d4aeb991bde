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


NOISE_DOMAINS = {'time': 0, 'bins': 1, 'lazy': 2, 'combined': 3}


def simo_sweep(engine, chan, snr_db, n_trials, num_rx, symbols_per_stream=14, seed=0, batch_trials=256,
               rank=0, world=1, noise_domain='lazy', fused=True, bits_host=None):
    """BER of the SIMO-MRC chain at every SNR point, `n_trials` independent streams per point.

    The default is the engine's fast path: spectral link at low Doppler and short delay spreads, fused channel +
    RX-FFT kernel otherwise, staged kernels where neither applies (the engine falls back by itself), with the AWGN added
    lazily by the consumers -- the same draws, hence bit-identical counts, as noise on the kept bins
    (`'bins'`).  `'combined'` draws one equivalent noise sample per MRC output instead of one per antenna
    (same BER statistics, different sample values, 1/R of the generator work): opt in by name.
    noise_domain: 'time' | 'bins' | 'lazy' | 'combined' (or LinkEngine.simo_ber's 0..3).
    bits_host: payload source in HOST memory instead of on-device random bits -- callable (trial_lo, n) ->
    uint8 [n * n_snr, ceil(nbits / 8)] of np.packbits() rows (ideally a pinned tensor); batches then go
    through LinkEngine.stream_host_batches (H2D of the next batch overlaps the current one)."""
    n_snr = len(snr_db)
    S, R = symbols_per_stream, num_rx
    nd = NOISE_DOMAINS[noise_domain] if isinstance(noise_domain, str) else int(noise_domain)
    snr_lin = torch.tensor([10 ** (s / 10) for s in snr_db], dtype=torch.float32, device=engine.device)
    nbits = S * engine.Nd * engine.bps
    if bits_host is not None:
        lo, hi = shard_range(n_trials, rank, world)
        counts = torch.zeros((2, n_snr), dtype=torch.int64)
        nb = min(batch_trials, max(hi - lo, 1))
        pipe = engine.stream_host_batches(chan, R, snr_lin.repeat(nb).repeat_interleave(R).contiguous(), nb * n_snr, S,
                                          nbits=nbits, seed=seed, noise_domain=nd)

        def batches():
            t = lo
            while t < hi:
                n = min(nb, hi - t)
                bits = bits_host(t, n)
                if n < nb:                                   # ragged last batch: pad with copies, count only the real rows
                    bits = torch.cat([bits, bits[:1].expand((nb - n) * n_snr, -1)])
                yield bits, t * n_snr, n
                t += n
        for err, n in pipe.run(batches()):
            counts[0] += err[:n * n_snr].view(n, n_snr).sum(dim=0)
            counts[1] += n * nbits
        counts = counts.to(engine.device)
        reduce_counts(counts)
        errors, bits = counts[0].cpu(), counts[1].cpu()
        return {'errors': errors, 'bits': bits, 'ber': errors.double() / bits.clamp(min=1).double()}
    # whole batches of `batch_trials` trials, two in flight where that pays (LinkEngine.simo_ber_batches); slot b of a batch always
    # runs at SNR point b % n_snr, so the per-slot counts accumulate in the workspaces and are reduced per point at
    # the end.  A ragged last batch follows on its own.
    lo, hi = shard_range(n_trials, rank, world)
    counts = torch.zeros((2, n_snr), dtype=torch.int64, device=engine.device)
    nfull = (hi - lo) // batch_trials
    if nfull > 0:
        B = batch_trials * n_snr
        nfl = min(engine.batches_in_flight(chan, B, R, S, fused=fused), nfull)
        wss = [engine.workspace(B, S, R, fading=chan.num_taps > 0, fused=fused, lazy=fused) for _ in range(nfl)]
        for w in wss:
            w['errors'].zero_()
        rows = snr_lin.repeat(batch_trials).repeat_interleave(R).contiguous()
        engine.simo_ber_batches(wss, chan, rows, seed, [(lo + k * batch_trials) * n_snr for k in range(nfull)],
                                noise_domain=nd, fused=fused)
        counts[0] += sum(w['errors'] for w in wss).view(batch_trials, n_snr).sum(dim=0)
        counts[1] += nfull * batch_trials * nbits
        del wss
    rest = (hi - lo) - nfull * batch_trials
    if rest > 0:
        t0 = lo + nfull * batch_trials
        ws = engine.workspace(rest * n_snr, S, R, fading=chan.num_taps > 0, fused=fused, lazy=fused)
        err = engine.simo_ber(ws, chan, snr_lin.repeat(rest).repeat_interleave(R).contiguous(), seed,
                              stream_id0=t0 * n_snr, noise_domain=nd, fused=fused)
        counts[0] += err.view(rest, n_snr).sum(dim=0)
        counts[1] += rest * nbits
    reduce_counts(counts)
    errors, bits = counts[0].cpu(), counts[1].cpu()
    return {'errors': errors, 'bits': bits, 'ber': errors.double() / bits.clamp(min=1).double()}


def scfdm_sweep(engine, chan, snr_db, n_trials, symbols_per_stream=14, seed=0, batch_trials=256, rank=0, world=1,
                papr_bins=200, papr_lo=0.0, papr_step=0.1, sc_fdm=True):
    """BASELINE config 2: BER + PAPR CCDF of the SISO SC-FDM uplink (reference simulate_siso with
    enable_sc_fdm=True, core/ofdm_core.py:660-737; PAPR per OFDM symbol as OFDMSystem.calculate_papr_without_cp,
    core/ofdm_system.py:173-229) in ONE pass per batch: the TX kernel's epilogue bins the PAPR of every OFDM
    symbol it produces.  Trials shard over ranks like simo_sweep; the histogram is all-reduced with the counters.
    Returns run_sweep's dict + 'papr_hist' [papr_bins], 'papr_edges_db', 'papr_ccdf' (P[PAPR > edge])."""
    n_snr = len(snr_db)
    S = symbols_per_stream
    snr_lin = torch.tensor([10 ** (s / 10) for s in snr_db], dtype=torch.float32, device=engine.device)
    hist = torch.zeros(papr_bins, dtype=torch.int64, device=engine.device)

    def count_batch(trial_lo, n):
        return engine.siso_ber(chan, snr_lin.repeat(n).contiguous(), S, seed, stream_id0=trial_lo * n_snr, sc_fdm=sc_fdm,
                               papr_hist=hist, papr_lo=papr_lo, papr_step=papr_step)

    out = run_sweep(count_batch, n_snr, n_trials, S * engine.Nd * engine.bps, batch_trials, rank, world, engine.device)
    reduce_counts(hist)
    h = hist.cpu()
    tot = max(int(h.sum()), 1)
    out['papr_hist'] = h
    out['papr_edges_db'] = papr_lo + papr_step * torch.arange(papr_bins + 1, dtype=torch.float64)
    out['papr_ccdf'] = 1.0 - torch.cumsum(h.double(), 0) / tot        # P[PAPR >= upper edge of bin i]
    return out


def simo_sweep_shared_channel(engine, chan, snr_db, n_trials, num_rx, symbols_per_stream=14, seed=0, batch_trials=4096,
                             rank=0, world=1, combine=True):
    """BER curve with the fading realisation of a trial shared by all SNR points (common random numbers along
    the SNR axis, the usual way to get smooth curves): per trial ONE pass of TX + fused channel + RX FFT gives the
    noise-free grid and the stream power; every SNR point then only runs the lazy-AWGN CRS estimate and the
    MRC / demap / count kernels on it, with its own noise draws.  Bits, Jakes phases and the channel are keyed by
    the global trial id, the noise by (SNR index, trial), so the counts do not depend on batching or sharding.
    Per SNR point this costs (TX + channel) / n_snr + CRS + MRC instead of the whole chain; the headline bench
    does NOT use it (there every (trial, SNR) pair is a fully independent link realisation)."""
    if chan.num_taps == 0:
        raise ValueError("a shared channel realisation only makes sense for a fading channel")
    n_snr = len(snr_db)
    S, R = symbols_per_stream, num_rx
    snr_lin = [float(10 ** (s / 10)) for s in snr_db]
    lo, hi = shard_range(n_trials, rank, world)
    counts = torch.zeros((2, n_snr), dtype=torch.int64, device=engine.device)
    nbits = S * engine.Nd * engine.bps
    from . import _native as nat
    t, ws = lo, None
    while t < hi:
        n = min(batch_trials, hi - t)
        if ws is None or ws['B'] != n:
            ws = engine.workspace(n, S, R, fading=True, fused=True)
        idx = engine.random_indices(n, S, seed, t, out=ws['idx'])
        tx, _, _ = engine.modulate(S, idx=idx, want_stats=False, out=ws['tx'])
        per = R * chan.num_taps * nat.LTE_JAKES_TONES
        ph = engine.random_phases(n, per, seed, t, out=ws['phases'].view(-1)[:n * per].view(n, per))
        got = engine.channel_rx_fft(tx, chan, n, R, S, ph, nat.WINDOW_USEFUL, out=ws['Y'], power=ws['power'])
        if got is None:
            raise ValueError("Doppler spread too large for the fused channel kernel; use simo_sweep")
        Y, power = got
        for si, sl in enumerate(snr_lin):
            rows = torch.full((n * R,), sl, dtype=torch.float32, device=engine.device)
            awgn = engine.awgn_desc(power, rows, seed, (si * n_trials + t) * R, combine=combine)
            H = engine.estimate(Y, n * R, S, nat.WINDOW_USEFUL, out=ws['H'], awgn=awgn)
            err = engine.mrc_demap_count(Y, H, idx, n, R, S, nbits=nbits, errors=ws['errors'], awgn=awgn)
            counts[0, si] += err.sum()
        counts[1] += n * nbits
        t += n
    reduce_counts(counts)
    errors, bits = counts[0].cpu(), counts[1].cpu()
    return {'errors': errors, 'bits': bits, 'ber': errors.double() / bits.clamp(min=1).double()}


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


def payload_sweep(config, bits, snr_range, n_iterations, modulations=('QPSK', '16-QAM', '64-QAM'),
                  num_rx_values=(1, 2, 4, 8), channel_type='rayleigh_mp', itu_profile='Pedestrian_A',
                  frequency_ghz=2.0, velocity_kmh=3.0, seed=0, rank=0, world=1, max_batch_bytes=4 << 30,
                  progress_callback=None, device=None):
    """The GUIs' full diversity sweep -- modulations x num_rx x SNR x iterations over one payload
    (reference SIMO/gui/main_window.py:128-273: four nested Python loops around simulate_siso /
    simulate_simo) -- as batched launches: for every (modulation, num_rx) all SNR points and
    iterations of a shard run as one batch of streams carrying the same payload bits, with
    independent Philox channel / noise draws keyed by the global stream id
    iteration * n_snr + snr_index.  num_rx = 1 uses the SISO zero-forcing chain, num_rx > 1 MRC.
    Returns the worker's result dict: {'mode': 'sweep_full', 'modulations', 'num_rx_values',
    'data': {mod: {'<n>RX': {'snr_values', 'ber_values', 'num_rx', 'errors', 'bits'}}}}."""
    import copy

    import numpy as np

    from .engine import LinkEngine, chan_for
    bits_t = torch.as_tensor(np.ascontiguousarray(np.asarray(bits).astype(np.uint8))) if not isinstance(bits, torch.Tensor) \
        else bits.to(torch.uint8)
    nbits = int(bits_t.numel())
    if nbits == 0:
        raise ValueError("Bits array cannot be empty")
    snr_range = [float(s) for s in np.atleast_1d(snr_range)]
    n_snr = len(snr_range)
    total = len(modulations) * len(num_rx_values)
    done = 0
    data = {}
    for mod in modulations:
        cfg = copy.copy(config)
        cfg.modulation = mod
        cfg._calculate_parameters()
        eng = LinkEngine.from_config(cfg, device=device)
        chan = chan_for(channel_type, cfg.fs, itu_profile, frequency_ghz, velocity_kmh)
        S = eng.symbols_for_bits(nbits)
        idx1 = eng.bits_to_indices(bits_t.to(eng.device).reshape(1, -1), nbits, S)
        snr_lin = torch.tensor([10 ** (s / 10) for s in snr_range], dtype=torch.float32, device=eng.device)
        data[mod] = {}
        for R in num_rx_values:
            # bytes per stream: tx + faded + Y + equalised symbols, with slack; whole iterations per batch
            per_stream = 8 * S * (eng.L * (1 + R) + eng.Nc * R + eng.Nd) + S * eng.Nd
            batch_trials = max(1, int(max_batch_bytes // (per_stream * n_snr)))
            state = {}

            def count_batch(trial_lo, n, R=R, state=state):
                B = n * n_snr
                if state.get('B') != B:
                    state['B'] = B
                    state['idx'] = idx1.expand(B, -1).contiguous()
                    state['rows'] = snr_lin.repeat(n).repeat_interleave(R).contiguous()
                    state['ws'] = eng.workspace(B, S, R, fading=chan.num_taps > 0, fused=True) if R > 1 else None
                if R == 1:
                    return eng.siso_ber(chan, state['rows'], S, seed, trial_lo * n_snr, idx=state['idx'], nbits=nbits)
                return eng.simo_ber(state['ws'], chan, state['rows'], seed, stream_id0=trial_lo * n_snr,
                                    idx=state['idx'], nbits=nbits, noise_domain=2, fused=True).clone()

            r = run_sweep(count_batch, n_snr, n_iterations, nbits, batch_trials, rank, world, eng.device)
            data[mod][f'{R}RX'] = {'snr_values': np.array(snr_range), 'ber_values': r['ber'].numpy(), 'num_rx': R,
                                   'errors': r['errors'].numpy(), 'bits': r['bits'].numpy()}
            state.clear()
            done += 1
            if progress_callback:
                progress_callback(int(15 + 80 * done / total), f"{mod} | {R}RX")
    return {'mode': 'sweep_full', 'modulations': list(modulations), 'num_rx_values': list(num_rx_values), 'data': data}


def coded_sweep(engine, chan, snr_db, n_trials, tb_bits, seed=0, batch_trials=64, rank=0, world=1, iterations=8):
    """BER and block-error rate of the coded SISO chain (reference OFDMSimulator.simulate_siso_coded,
    core/ofdm_core.py:925-1338) at every SNR point: every stream carries its own random transport
    block of `tb_bits` bits (Philox keyed by the global stream id), CRC-24A, turbo code, interleaver,
    channel, ZF, max-log LLRs and `iterations` max-log BCJR iterations.  Adds 'bler' (fraction of
    transport blocks whose CRC-24A failed) to run_sweep's result."""
    n_snr = len(snr_db)
    snr_t = torch.tensor([float(s) for s in snr_db], dtype=torch.float32, device=engine.device)
    fails = torch.zeros((2, n_snr), dtype=torch.int64, device=engine.device)
    S_src = -(-tb_bits // engine.Nd)

    def count_batch(trial_lo, n):
        B = n * n_snr
        sid0 = trial_lo * n_snr
        bits = (engine.random_indices(B, S_src, seed, sid0)[:, :tb_bits] & 1).contiguous()
        err, crc_ok = engine.siso_coded_ber(bits, chan, snr_t.repeat(n).contiguous(), seed, sid0, iterations)
        fails[0] += (1 - crc_ok.long()).view(n, n_snr).sum(dim=0)
        fails[1] += n
        return err

    out = run_sweep(count_batch, n_snr, n_trials, tb_bits, batch_trials, rank, world, engine.device)
    reduce_counts(fails)
    out['block_errors'] = fails[0].cpu()
    out['bler'] = fails[0].double().cpu() / fails[1].clamp(min=1).double().cpu()
    return out


def sfbc_sweep(config, snr_db, n_trials, num_rx=2, itu_profile='Pedestrian_A', frequency_ghz=2.0, velocity_kmh=3.0,
               symbols_per_stream=14, seed=0, batch_trials=128, rank=0, world=1, device=None):
    """BASELINE config 4: Monte-Carlo BER of 2 x num_rx SFBC-Alamouti transmit diversity over Rayleigh
    multipath (reference simulate_miso / simulate_mimo), trials sharded over ranks like simo_sweep."""
    from .engine import LinkEngine, chan_for
    from . import tables
    eng0 = LinkEngine.from_config(config, device=device)
    eng = LinkEngine.from_config(config, pilot_sets=tables.mimo_pilot_sets(2, eng0.Np), device=device)
    chan = chan_for('rayleigh_mp', config.fs, itu_profile, frequency_ghz, velocity_kmh)
    n_snr = len(snr_db)
    S, R = symbols_per_stream, num_rx
    snr_lin = torch.tensor([10 ** (s / 10) for s in snr_db], dtype=torch.float32, device=eng.device)

    def count_batch(trial_lo, n):
        rows = snr_lin.repeat(n).repeat_interleave(R).contiguous()
        return eng.sfbc_ber(chan, rows, S, R, seed, stream_id0=trial_lo * n_snr)

    return run_sweep(count_batch, n_snr, n_trials, S * 2 * (eng.Nd // 2) * eng.bps, batch_trials, rank, world,
                     eng.device)


def sm_sweep(config, snr_db, n_trials, num_tx=4, num_rx=4, rank=4, detector='MMSE', itu_profile='Pedestrian_A',
             frequency_ghz=2.0, velocity_kmh=3.0, symbols_per_stream=1, seed=0, batch_trials=64, rank_id=0, world=1,
             device=None, precoder=None, feedback=None, feedback_block=16):
    """BASELINE config 5: Monte-Carlo BER of T x R spatial multiplexing with codebook precoding and an
    MMSE / ZF / SIC / MRC detector (reference simulate_spatial_multiplexing), trials sharded over ranks.

    With a fixed rank all SNR points of a run of trials share one pass (the detector takes sigma^2 per stream);
    stream ids are trial-major there, id = trial * n_snr + snr_index, so a pass is one contiguous id range.
    With adaptive rank the precoder changes with the SNR point and the feedback block, so a pass is (one SNR
    point, one block) and ids are SNR-major: id = snr_index * n_trials + trial.
    rank 1..4: `precoder(rank) -> W [T, rank]` (codebook entry 0, the reference's fixed-rank branch).
    rank 'adaptive': the reference draws an H_initial unrelated to the channel and feeds it to
    RankAdaptation (core/ofdm_core.py:2573-2583); here one H_initial is drawn per (SNR point, block of
    `feedback_block` consecutive global trials) -- keyed by (seed, snr_index, block), so sharding cannot
    change it.  By default (feedback=None) the draws (lte_random_channel) and the reference's rule -- RI from
    the eigenvalues of H^H H with the SNR gates, PMI by the capacity metric (lte_rank_feedback) -- run on the
    device, all (SNR point, block) units of a batch in one launch each; a callable
    `feedback(H, snr_db) -> (ri, pmi, W)` keeps the rule (and NumPy draws) on the host.
    Adds 'rank_hist' [n_snr, 4] (streams per rank) to run_sweep's result."""
    import numpy as np

    from . import tables
    from .engine import LinkEngine, chan_for
    eng0 = LinkEngine.from_config(config, device=device)
    eng = LinkEngine.from_config(config, pilot_sets=tables.mimo_pilot_sets(num_tx, eng0.Np), device=device)
    chan = chan_for('rayleigh_mp', config.fs, itu_profile, frequency_ghz, velocity_kmh, gain_conversions=3)
    n_snr = len(snr_db)
    S, R = symbols_per_stream, num_rx
    adaptive = rank == 'adaptive'
    if not adaptive and precoder is None:
        raise ValueError("a fixed rank needs precoder(rank) -> W")
    rank_hist = torch.zeros((n_snr, 4), dtype=torch.int64, device=eng.device)
    device_feedback = adaptive and feedback is None
    if device_feedback:
        from core.codebook_lte import LTECodebook
        max_rank = min(num_tx, num_rx, 4)
        cb_tab, cb_sizes = eng.rank_codebook(num_tx, max_rank)
        books = [LTECodebook(num_tx, transmission_mode='TM4', rank=r) for r in range(1, max_rank + 1)]
        n_blocks = -(-n_trials // feedback_block)
        snr_t = torch.tensor([float(x) for x in snr_db], dtype=torch.float64, device=eng.device)

    def device_units(trial_lo, n):
        """(ri, pmi) of every (SNR point, feedback block) unit the trials [trial_lo, trial_lo + n) touch."""
        blk0, blk1 = trial_lo // feedback_block, (trial_lo + n - 1) // feedback_block
        nb = blk1 - blk0 + 1
        H = torch.cat([eng.random_channel(nb, num_rx, num_tx, seed ^ 0x5249, stream_id0=si * n_blocks + blk0)
                       for si in range(n_snr)]) * (1.0 / np.sqrt(num_tx))
        ri, pmi = eng.rank_feedback(H, snr_t.repeat_interleave(nb).contiguous(), cb_tab, cb_sizes, max_rank=max_rank)
        return blk0, ri.view(n_snr, nb).cpu().numpy(), pmi.view(n_snr, nb).cpu().numpy()

    def count_batch(trial_lo, n):
        if not adaptive:
            # all SNR points in one pass: stream j of the pass is (trial trial_lo + j // n_snr, snr j % n_snr)
            snr_streams = torch.tensor(snr_db, dtype=torch.float64, device=eng.device).repeat(n)
            err = eng.sm_ber(chan, precoder(int(rank)), snr_streams, n * n_snr, S, R, detector, seed,
                             stream_id0=trial_lo * n_snr)
            rank_hist[:, int(rank) - 1] += n
            return err.reshape(-1)
        err = torch.zeros((n, n_snr), dtype=torch.int64, device=eng.device)
        if device_feedback:
            blk0, ri_u, pmi_u = device_units(trial_lo, n)
        for si, snr in enumerate(snr_db):
            t = trial_lo
            while t < trial_lo + n:
                if device_feedback:
                    blk = t // feedback_block
                    end = min(trial_lo + n, (blk + 1) * feedback_block)
                    ri = int(ri_u[si, blk - blk0])
                    W = books[ri - 1].get_precoder(int(pmi_u[si, blk - blk0]))
                elif adaptive:
                    blk = t // feedback_block
                    end = min(trial_lo + n, (blk + 1) * feedback_block)
                    rs = np.random.RandomState([seed & 0x7fffffff, si, blk])
                    Hi = (rs.randn(num_rx, num_tx) + 1j * rs.randn(num_rx, num_tx)) / np.sqrt(2 * num_tx)
                    ri, _, W = feedback(Hi, float(snr))
                else:
                    end, ri, W = trial_lo + n, int(rank), precoder(int(rank))
                m = end - t
                err[t - trial_lo:end - trial_lo, si] = eng.sm_ber(chan, W, float(snr), m, S, R, detector, seed,
                                                                  stream_id0=si * n_trials + t)
                rank_hist[si, ri - 1] += m
                t = end
        return err.reshape(-1)

    out = run_sweep(count_batch, n_snr, n_trials, S * eng.Nd * eng.bps, batch_trials, rank_id, world, eng.device)
    reduce_counts(rank_hist)
    out['rank_hist'] = rank_hist.cpu()
    return out
