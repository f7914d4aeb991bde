"""Sweep engine on the GPU: counts are independent of the batch size and of the trial sharding,
and the BER curve behaves (monotone in SNR, diversity gain with more antennas)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(bw=1.25, mod='16-QAM', prof='Pedestrian_A', v=3.0):
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(bw, 15.0, mod)
    return LinkEngine.from_config(cfg), chan_for('rayleigh_mp', cfg.fs, prof, 2.0, v)


def test_counts_do_not_depend_on_batching_or_sharding():
    from lte_b200.sweep import simo_sweep
    eng, chan = _setup()
    snr = [0.0, 10.0, 20.0]
    a = simo_sweep(eng, chan, snr, n_trials=40, num_rx=2, seed=5, batch_trials=40)
    b = simo_sweep(eng, chan, snr, n_trials=40, num_rx=2, seed=5, batch_trials=7)
    assert torch.equal(a['errors'], b['errors'])
    # two "ranks" run one after the other on this GPU and summed by hand == one rank
    parts = [simo_sweep(eng, chan, snr, n_trials=40, num_rx=2, seed=5, batch_trials=16, rank=r, world=2)
             for r in range(2)]
    assert torch.equal(parts[0]['errors'] + parts[1]['errors'], a['errors'])
    assert torch.equal(parts[0]['bits'] + parts[1]['bits'], a['bits'])
    c = simo_sweep(eng, chan, snr, n_trials=40, num_rx=2, seed=6, batch_trials=40)
    assert not torch.equal(a['errors'], c['errors'])


def test_ber_curve_shape_and_diversity_gain():
    from lte_b200.sweep import simo_sweep
    eng, chan = _setup()
    snr = [0.0, 8.0, 16.0, 24.0]
    r1 = simo_sweep(eng, chan, snr, n_trials=200, num_rx=1, seed=1)
    r4 = simo_sweep(eng, chan, snr, n_trials=200, num_rx=4, seed=1)
    b1, b4 = r1['ber'].numpy(), r4['ber'].numpy()
    assert np.all(np.diff(b1) < 0) and np.all(np.diff(b4) < 0)
    assert np.all(b4 < b1)
    assert 0.2 < b1[0] < 0.5


def test_time_and_frequency_domain_noise_agree_statistically():
    from lte_b200.sweep import simo_sweep
    eng, chan = _setup()
    snr = [6.0, 12.0]
    t = simo_sweep(eng, chan, snr, n_trials=400, num_rx=2, seed=2, noise_domain=0)['ber'].numpy()
    f = simo_sweep(eng, chan, snr, n_trials=400, num_rx=2, seed=2, noise_domain=1)['ber'].numpy()
    assert np.all(np.abs(t - f) / t < 0.05)


@pytest.mark.parametrize('bw,mod,R', [(1.25, '16-QAM', 2), (5.0, '64-QAM', 4), (2.5, 'QPSK', 1)])
def test_lazy_awgn_is_bit_identical_to_noise_in_the_rx_epilogue(bw, mod, R):
    """noise_domain 2 keeps Y noise-free and lets the CRS / MRC kernels add the very same Philox
    draws while reading it: every stream's error count must match noise_domain 1 exactly."""
    from lte_b200 import _native as nat
    eng, chan = _setup(bw=bw, mod=mod)
    B, S = 24, 15                                   # two slots, the second one symbol long
    ws = eng.workspace(B, S, R, fading=True)
    snr = torch.tensor([10 ** (s / 10) for s in (2.0, 9.0, 17.0)], dtype=torch.float32, device='cuda')
    rows = snr.repeat(B // 3).repeat_interleave(R).contiguous()
    e1 = eng.simo_ber(ws, chan, rows, seed=9, stream_id0=7, noise_domain=1).clone()
    Y1, H1 = ws['Y'].clone(), ws['H'].clone()
    e2 = eng.simo_ber(ws, chan, rows, seed=9, stream_id0=7, noise_domain=2).clone()
    assert torch.equal(e1, e2) and int(e1.sum()) > 0
    # the channel estimate formed from the lazily-noised pilots is the same tensor, bit for bit
    assert torch.equal(ws['H'], H1)
    assert not torch.equal(ws['Y'], Y1)             # ...while Y itself stayed noise-free
    # AWGN channel type (no fading buffer, rx_div = R)
    from lte_b200 import chan_for
    awgn = chan_for('awgn', eng.fs, 'Pedestrian_A', 2.0, 0.0)
    wa = eng.workspace(B, S, R, fading=False)
    a1 = eng.simo_ber(wa, awgn, rows, seed=3, noise_domain=1).clone()
    a2 = eng.simo_ber(wa, awgn, rows, seed=3, noise_domain=2).clone()
    assert torch.equal(a1, a2)


def test_combined_mrc_noise_and_fused_path_agree_statistically():
    """noise_domain 3 draws one equivalent sample per MRC output; the fused channel+FFT kernel changes
    only fp32 rounding.  Both must reproduce the BER curve of the per-antenna RX-epilogue noise."""
    from lte_b200.sweep import simo_sweep
    eng, chan = _setup(bw=2.5, mod='16-QAM')
    snr = [4.0, 10.0, 16.0]
    base = simo_sweep(eng, chan, snr, n_trials=600, num_rx=4, seed=3, noise_domain=1)
    comb = simo_sweep(eng, chan, snr, n_trials=600, num_rx=4, seed=3, noise_domain=3)
    fused = simo_sweep(eng, chan, snr, n_trials=600, num_rx=4, seed=3, noise_domain=3, fused=True)
    same = simo_sweep(eng, chan, snr, n_trials=600, num_rx=4, seed=3, noise_domain=2, fused=True)
    b0 = base['ber'].numpy()
    assert not torch.equal(base['errors'], comb['errors'])          # different draws ...
    assert np.all(np.abs(comb['ber'].numpy() - b0) / b0 < 0.05)     # ... same statistics
    assert np.all(np.abs(fused['ber'].numpy() - b0) / b0 < 0.05)
    # per-antenna draws through the fused kernel: only slicer-boundary flips separate it from the staged path
    assert np.all(np.abs(same['errors'].numpy() - base['errors'].numpy()) <= np.maximum(3, base['errors'].numpy() // 2000))


def test_shared_channel_sweep_matches_the_independent_sweep_statistically():
    """Common random numbers along the SNR axis: same BER curve as fully independent streams (within Monte-Carlo
    error), invariant to batching and sharding, and the first SNR point of a 1-point sweep reproduces exactly."""
    from lte_b200.sweep import simo_sweep, simo_sweep_shared_channel
    eng, chan = _setup(bw=1.25, mod='16-QAM')
    snr = [2.0, 8.0, 14.0, 20.0]
    a = simo_sweep_shared_channel(eng, chan, snr, 600, 2, seed=3, batch_trials=600)
    b = simo_sweep_shared_channel(eng, chan, snr, 600, 2, seed=3, batch_trials=77)
    assert torch.equal(a['errors'], b['errors'])
    parts = [simo_sweep_shared_channel(eng, chan, snr, 600, 2, seed=3, batch_trials=128, rank=r, world=3) for r in range(3)]
    assert torch.equal(sum(p['errors'] for p in parts), a['errors'])
    assert torch.equal(a['bits'], torch.full((4,), 600 * 14 * eng.Nd * eng.bps, dtype=torch.int64))
    ind = simo_sweep(eng, chan, snr, n_trials=600, num_rx=2, seed=4, noise_domain=3, fused=True)
    ba, bi = a['ber'].numpy(), ind['ber'].numpy()
    assert np.all(np.diff(ba) < 0) and np.all(np.abs(ba - bi) / bi < 0.2)
    awgn_only = __import__('lte_b200').chan_for('awgn', eng.fs, 'Pedestrian_A', 2.0, 0.0)
    with pytest.raises(ValueError):
        simo_sweep_shared_channel(eng, awgn_only, snr, 4, 2)


def test_simo_ber_accumulate_equals_the_sum_of_separate_passes():
    """simo_ber(accumulate=True) adds into the workspace's per-slot counters (what bench.py's timed loop does):
    three passes accumulate to the sum of three separate passes, on the spectral, fused and staged paths."""
    import torch
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(2.5, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg)
    B, S, R = 6, 14, 2
    rows = torch.tensor([3.0, 30.0], dtype=torch.float32, device='cuda').repeat(B // 2).repeat_interleave(R).contiguous()
    for prof, v, kw in (('Pedestrian_A', 3.0, dict(fused=True, noise_domain=3)), ('Vehicular_A', 60.0, dict(fused=True, noise_domain=2)),
                        ('Pedestrian_B', 3.0, dict(fused=False, noise_domain=1))):
        chan = chan_for('rayleigh_mp', cfg.fs, prof, 2.0, v)
        ws = eng.workspace(B, S, R, fading=True, fused=kw['fused'], lazy=kw['fused'])
        want = sum(eng.simo_ber(ws, chan, rows, 5, stream_id0=i * B, **kw).clone() for i in range(3))
        ws['errors'].zero_()
        for i in range(3):
            got = eng.simo_ber(ws, chan, rows, 5, stream_id0=i * B, accumulate=True, **kw)
        assert torch.equal(got, want) and int(want.sum()) > 0


@pytest.mark.parametrize('bw,mod,R,B', [(20.0, '64-QAM', 4, 320), (5.0, '16-QAM', 2, 48)])
def test_batches_in_flight_count_like_batches_in_sequence(bw, mod, R, B):
    """LinkEngine.simo_ber_batches (what bench.py times and simo_sweep runs): seven batches alternating between two
    workspaces on two streams accumulate exactly the per-slot counts of the same batches one after the other, and the
    engine asks for two batches in flight only where the spectral link runs."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(bw, 15.0, mod)
    eng = LinkEngine.from_config(cfg)
    S = 14
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    rows = torch.tensor([6.0, 40.0, 300.0, 2000.0], dtype=torch.float32, device='cuda').repeat(B // 4).repeat_interleave(R).contiguous()
    sids = [k * B for k in range(7)]
    one = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    want = sum(eng.simo_ber(one, chan, rows, 8, stream_id0=s, fused=True, noise_domain=3).clone() for s in sids)
    assert eng.batches_in_flight(chan, B, R, S) == 2
    assert eng.batches_in_flight(chan_for('rayleigh_mp', cfg.fs, 'Vehicular_A', 2.0, 60.0), B, R, S) == 1
    wss = [eng.workspace(B, S, R, fading=True, fused=True, lazy=True) for _ in range(2)]
    for _ in range(2):                                          # twice: the accumulators restart from zero
        for w in wss:
            w['errors'].zero_()
        eng.simo_ber_batches(wss, chan, rows, 8, sids, noise_domain=3)
        got = wss[0]['errors'] + wss[1]['errors']               # on the current stream: simo_ber_batches has joined it
        assert torch.equal(got, want) and int(want.sum()) > 0
