"""Spectral fading link (lte_tx_spectral + lte_channel_spectral): checked against the oracle's
sample-by-sample path, against the fp64 restatement of its own algebra (tests/spectral_ref.py), against the
fused time-domain kernel it stands in for, and through the error counts of the sweep built on it."""
import numpy as np
import pytest
import torch

import spectral_ref as SR
from gpu_chain import to_dev
from helpers import rel_err
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu

CASES = [(20.0, '64-QAM', 'Pedestrian_A', 4, 14, 3.0), (5.0, 'QPSK', 'Pedestrian_A', 2, 15, 3.0),
         (2.5, '16-QAM', 'Pedestrian_B', 3, 5, 3.0), (10.0, '16-QAM', 'Vehicular_A', 1, 3, 3.0),
         (2.5, '64-QAM', 'Vehicular_B', 8, 2, 2.0), (20.0, '16-QAM', 'Bad_Urban', 5, 2, 1.0)]


def _unpair(Yc, n):
    """Rows of the compact layout are padded to an even length: [..., 2 pairs] -> [..., n]."""
    return Yc[..., :n]


def _setup(bw, mod, prof, R, S, v, B=2, seed=5):
    from lte_b200 import LinkEngine, chan_for
    num = O.Numerology(bw, 15.0, mod)
    eng = LinkEngine(num.N, num.Nc, num.cp_length, num.bits_per_symbol, num.fs)
    rs = np.random.RandomState(seed)
    nbits = eng.Nd * eng.bps * S
    bits = rs.randint(0, 2, (B, nbits))
    idx = eng.bits_to_indices(to_dev(bits.astype(np.uint8), torch.uint8), nbits, S)
    chan = chan_for('rayleigh_mp', num.fs, prof, 2.0, v)
    d, g = O.itu_taps(prof, num.fs)
    ph = 2 * np.pi * rs.rand(B, R, len(d), 16)
    return num, eng, bits, idx, chan, d, g, ph


@pytest.mark.parametrize('bw,mod,prof,R,S,v', CASES)
def test_spectral_grid_and_power_match_oracle(bw, mod, prof, R, S, v):
    from lte_b200 import _native as nat
    num, eng, bits, idx, chan, d, g, ph = _setup(bw, mod, prof, R, S, v)
    B = bits.shape[0]
    G, tail = eng.tx_spectral(S, idx)
    u = to_dev(ph / (2 * np.pi), torch.float32)
    got = eng.channel_spectral(idx, G, tail, chan, B, R, S, u)
    assert got is not None
    Y, power = got
    Yn, pw = Y.cpu().numpy().reshape(B, R, S, -1), power.cpu().numpy()
    k0, nk = eng.window(nat.WINDOW_USEFUL)
    kept = np.arange(k0, k0 + nk)
    fD = O.doppler_hz(2.0, v)
    Gn, tn = G.cpu().numpy().reshape(B, S, nk), tail.cpu().numpy().reshape(B, S, -1)
    for b in range(B):
        sig, _ = O.modulate_stream(bits[b], num)
        x = sig.reshape(S, num.L)
        # TX side: symbol tails and the ramp-weighted spectrum
        assert rel_err(tn[b], x[:, num.N:]) < 2e-6
        nc = 0.5 * (num.L - 1) - num.cp_length
        Gw = np.fft.fft((np.arange(num.N) - nc)[None, :] * x[:, num.cp_length:], axis=-1) / np.sqrt(num.N)
        assert rel_err(Gn[b], Gw[:, kept]) < 2e-6
        # channel side vs the oracle's time-domain path (budget 1e-5) and vs its own algebra in fp64 (tighter)
        Yt, Pt = SR.time_domain_rx(sig, num, fD, list(d), g, ph[b])
        Ys, Ps = SR.spectral_rx(sig, num, fD, list(d), g, ph[b], kept)
        for r in range(R):
            assert rel_err(Yn[b, r], Yt[r][:, kept]) < 1e-5
            assert rel_err(Yn[b, r], Ys[r][:, kept]) < 2e-6
            assert abs(pw[b, r] / Pt[r] - 1) < 1e-5
            assert abs(pw[b, r] / Ps[r] - 1) < 2e-6
        assert np.abs(Yn[b] - Yt[:, :, kept]).max() / np.abs(Yt[:, :, kept]).max() < 1e-5


@pytest.mark.parametrize('bw,mod,prof,R,S,v', CASES[:4])
def test_spectral_matches_fused_time_domain_kernel(bw, mod, prof, R, S, v):
    from lte_b200 import _native as nat
    num, eng, bits, idx, chan, d, g, ph = _setup(bw, mod, prof, R, S, v, B=3)
    B = bits.shape[0]
    u = to_dev(ph / (2 * np.pi), torch.float32)
    G, tail = eng.tx_spectral(S, idx)
    Y, power = eng.channel_spectral(idx, G, tail, chan, B, R, S, u)
    tx, _, _ = eng.modulate(S, idx=idx, want_stats=False)
    Yf, pf = eng.channel_rx_fft(tx, chan, B, R, S, u, nat.WINDOW_USEFUL)
    assert rel_err(Y.cpu().numpy(), Yf.cpu().numpy()) < 2e-6
    assert np.allclose(power.cpu().numpy(), pf.cpu().numpy(), rtol=2e-6)


@pytest.mark.parametrize('bw,mod,prof,R,S,v', [CASES[0], CASES[1], (5.0, '16-QAM', 'Pedestrian_A', 3, 29, 3.0),
                                              (10.0, '16-QAM', 'Pedestrian_A', 4, 14, 3.0),
                                              (2.5, 'QPSK', 'Pedestrian_A', 2, 15, 3.0)])
def test_compact_layout_is_a_gather_of_the_window(bw, mod, prof, R, S, v):
    """COMPACT output = data bins in data-symbol order + the pilot bins of every slot's first symbol."""
    from lte_b200 import _native as nat
    num, eng, bits, idx, chan, d, g, ph = _setup(bw, mod, prof, R, S, v, B=3)
    B = bits.shape[0]
    u = to_dev(ph / (2 * np.pi), torch.float32)
    G, tail = eng.tx_spectral(S, idx)
    Y, power = eng.channel_spectral(idx, G, tail, chan, B, R, S, u)
    (Yd, Yp), power_c = eng.channel_spectral(idx, G, tail, chan, B, R, S, u, compact=True)
    k0, nk = eng.window(nat.WINDOW_USEFUL)
    di = torch.as_tensor(eng.data_idx - k0, device='cuda')
    pi = torch.as_tensor(eng.pilot_idx - k0, device='cuda')
    assert torch.equal(_unpair(Yd, eng.Nd), Y[:, :, di])
    assert torch.equal(_unpair(Yp, eng.Np), Y[:, ::14, :][:, :, pi])
    assert torch.equal(power, power_c)


@pytest.mark.parametrize('combine', [False, True])
@pytest.mark.parametrize('bw,mod,prof,R,S,v', [CASES[0], (5.0, '16-QAM', 'Pedestrian_A', 3, 29, 3.0)])
def test_compact_consumers_count_exactly_like_the_windowed_ones(bw, mod, prof, R, S, v, combine):
    """lte_crs_ls_compact + lte_mrc_demap_count_compact vs lte_crs_ls_interp_awgn + lte_mrc_demap_count_awgn on the
    same noise-free grid and the same lazy AWGN draws: per-stream error counts must be identical."""
    from lte_b200 import _native as nat
    num, eng, bits, idx, chan, d, g, ph = _setup(bw, mod, prof, R, S, v, B=6)
    B = bits.shape[0]
    u = to_dev(ph / (2 * np.pi), torch.float32)
    G, tail = eng.tx_spectral(S, idx)
    Y, power = eng.channel_spectral(idx, G, tail, chan, B, R, S, u)
    (Yd, Yp), _ = eng.channel_spectral(idx, G, tail, chan, B, R, S, u, compact=True)
    snr = torch.tensor([10 ** (x / 10) for x in (2.0, 8.0, 14.0, 20.0, 26.0, 40.0)], dtype=torch.float32, device='cuda')
    rows = snr.repeat_interleave(R).contiguous()
    nbits = S * eng.Nd * eng.bps - 5
    for awgn in (None, eng.awgn_desc(power, rows, 9, 77, combine=combine)):
        H = eng.estimate(Y, B * R, S, nat.WINDOW_USEFUL, awgn=awgn)
        want = eng.mrc_demap_count(Y, H, idx, B, R, S, nbits=nbits, awgn=awgn).clone()
        Hp = eng.estimate_compact(Yp, B * R, S, awgn=awgn)
        got = eng.mrc_demap_count_compact(Yd, Hp, idx, B, R, S, nbits=nbits, awgn=awgn)
        # ... and with the LS estimate formed inside the MRC launch (what the sweep runs)
        one = eng.mrc_demap_count_compact(Yd, None, idx, B, R, S, nbits=nbits, awgn=awgn, Yp=Yp)
        assert torch.equal(one, got)
        assert torch.equal(got, want)
        if awgn is not None:
            assert int(want.sum()) > 0
        k0, nk = eng.window(nat.WINDOW_USEFUL)
        pi = torch.as_tensor(eng.pilot_idx - k0, device='cuda')
        assert torch.equal(Hp, H[:, :, pi])


@pytest.mark.parametrize('bw,mod,R,prof', [(2.5, '16-QAM', 2, 'Pedestrian_A'), (5.0, '64-QAM', 4, 'Vehicular_A'),
                                           (2.5, 'QPSK', 1, 'Pedestrian_B'), (20.0, '64-QAM', 4, 'Pedestrian_A')])
def test_spectral_sweep_counts_track_the_fused_path(bw, mod, R, prof):
    """Same draws, same fading polynomial, different evaluation (bin by bin instead of sample by sample): per-stream
    error counts may differ only through symbols that sit on a slicer boundary."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(bw, 15.0, mod)
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, prof, 2.0, 3.0)
    B, S = 12, 15
    wf = eng.workspace(B, S, R, fading=True, fused=True)
    wsp = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    snr = torch.tensor([10 ** (s / 10) for s in (4.0, 12.0, 22.0)], dtype=torch.float32, device='cuda')
    rows = snr.repeat(B // 3).repeat_interleave(R).contiguous()
    for nd in (2, 3):
        e_fused = eng.simo_ber(wf, chan, rows, seed=4, stream_id0=11, fused=True, spectral=False, noise_domain=nd).clone()
        e_spec = eng.simo_ber(wsp, chan, rows, seed=4, stream_id0=11, fused=True, spectral=True, noise_domain=nd).clone()
        assert wsp['spectral'] and 'tx' not in wsp                  # the spectral link really ran
        bits = S * eng.Nd * eng.bps
        assert int(e_fused.sum()) > 0
        assert int((e_fused - e_spec).abs().max()) <= max(2, bits // 20000)
        assert abs(int(e_fused.sum()) - int(e_spec.sum())) <= max(3, int(e_fused.sum()) // 2000)


@pytest.mark.parametrize('bw,mod,prof,R,S', [(20.0, '64-QAM', 'Pedestrian_A', 4, 14), (5.0, '16-QAM', 'Pedestrian_A', 2, 29),
                                            (2.5, 'QPSK', 'Pedestrian_B', 3, 15)])
def test_persistent_walk_does_not_depend_on_the_batch(bw, mod, prof, R, S):
    """The channel kernel is persistent: with more streams than SMs a CTA walks several streams one after the other
    (the next stream's first symbol has the previous stream's samples in front of it in the tail array, the power
    and slot counters restart, the ring never drains).  A stream's grid and power must not depend on which other
    streams share its launch: 400 streams in one launch against the same streams in launches of at most 100."""
    num, eng, bits, idx, chan, d, g, ph = _setup(bw, mod, prof, R, S, 3.0, B=400, seed=9)
    B = bits.shape[0]
    u = to_dev(ph / (2 * np.pi), torch.float32)
    G, tail = eng.tx_spectral(S, idx)
    (Yd, Yp), power = eng.channel_spectral(idx, G, tail, chan, B, R, S, u, compact=True)
    Yd, Yp, power = Yd.clone(), Yp.clone(), power.clone()
    assert torch.isfinite(torch.view_as_real(Yd)).all() and float(power.min()) > 0
    for lo, hi in [(0, 100), (100, 149), (149, 300), (300, 400)]:
        n = hi - lo
        Gs, ts = eng.tx_spectral(S, idx[lo:hi].contiguous())
        assert torch.equal(Gs, G[lo * S:hi * S]) and torch.equal(ts, tail[lo * S:hi * S])
        (yd, yp), pw = eng.channel_spectral(idx[lo:hi].contiguous(), Gs, ts, chan, n, R, S, u[lo:hi].contiguous(), compact=True)
        assert torch.equal(yd, Yd[lo * R:hi * R])
        assert torch.equal(yp, Yp[lo * R:hi * R])
        assert torch.equal(pw, power[lo:hi])


def test_bench_size_batch_without_noise_decodes_the_streams():
    """BASELINE's full size (20 MHz, 64-QAM, 1x4, 4096 subframes per batch: what bench.py times) through the sweep's
    default pipeline at 100 dB (size-independent property: encode -> channel -> decode round trip).  What is left
    without noise is the reference estimator's own floor -- one LS estimate per 14 symbols, linear between pilots,
    at 3 km/h -- a few bits in 3.4e8: the BER must stay below 1e-6 with more than 99.5 % of the streams clean, for
    the first batch of stream ids and the next."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    B, S, R = 4096, 14, 4
    ws = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    rows = torch.full((B * R,), 1e10, dtype=torch.float32, device='cuda')
    for k in range(2):
        err = eng.simo_ber(ws, chan, rows, seed=3, stream_id0=k * B, fused=True, noise_domain=3)
        assert ws.get('spectral') is True
        assert int(err.sum()) <= 1e-6 * B * S * eng.Nd * eng.bps
        assert int((err > 0).sum()) <= B // 200
        assert float(ws['power'].min()) > 0
    # and the same streams at 10 dB do see errors, stream by stream the same whether they run alone or in the batch
    rows = torch.full((B * R,), 10.0, dtype=torch.float32, device='cuda')
    big = eng.simo_ber(ws, chan, rows, seed=3, stream_id0=0, fused=True, noise_domain=3).clone()
    assert int(big.sum()) > 0
    w2 = eng.workspace(64, S, R, fading=True, fused=True, lazy=True)
    small = eng.simo_ber(w2, chan, rows[:64 * R].contiguous(), seed=3, stream_id0=1000, fused=True, noise_domain=3)
    assert torch.equal(small, big[1000:1064])


def test_spectral_reports_unsupported():
    from lte_b200 import chan_for
    num, eng, bits, idx, chan, d, g, ph = _setup(5.0, 'QPSK', 'Pedestrian_A', 2, 2, 3.0)
    assert eng.spectral_workspace_bytes(chan_for('awgn', num.fs), 2, 2, 2) is None
    assert eng.spectral_workspace_bytes(chan_for('rayleigh_mp', num.fs, 'Pedestrian_A', 2.0, 30.0), 2, 2, 2) is None
    assert eng.spectral_workspace_bytes(chan_for('rayleigh_mp', num.fs, 'Pedestrian_A', 2.0, 3.0), 2, 2, 2) > 0
    num2 = O.Numerology(1.25, 15.0, 'QPSK')
    from lte_b200 import LinkEngine
    eng2 = LinkEngine(num2.N, num2.Nc, 4, 2, num2.fs)           # a 4-sample prefix cannot hold Vehicular_B's delay spread
    assert eng2.spectral_workspace_bytes(chan_for('rayleigh_mp', num2.fs, 'Vehicular_B', 2.0, 1.0), 1, 1, 1) is None
    eng3 = LinkEngine(num2.N, num2.Nc, num2.cp_length, 2, num2.fs)   # cp = 9: rows of the tail array are not 16-byte runs
    assert eng3.spectral_workspace_bytes(chan_for('rayleigh_mp', num2.fs, 'Pedestrian_A', 2.0, 3.0), 1, 1, 1) is None


def test_automatic_choice_leaves_long_delay_spreads_to_the_time_domain_kernel():
    """simo_ber picks the spectral link by itself only while the Horner sweep is short (<= SPECTRAL_MAX_DELAY samples):
    Vehicular_A at 20 MHz (77 samples) goes through the fused time-domain kernel unless spectral=True asks for it;
    both give the same counts up to slicer-boundary symbols."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(20.0, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg)
    B, S, R = 4, 14, 2
    rows = torch.full((B * R,), 10.0, dtype=torch.float32, device='cuda')
    short = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    long_ = chan_for('rayleigh_mp', cfg.fs, 'Vehicular_A', 2.0, 3.0)
    assert max(short.delay[:short.num_taps]) <= eng.SPECTRAL_MAX_DELAY < max(long_.delay[:long_.num_taps])
    w1 = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    eng.simo_ber(w1, short, rows, seed=1, fused=True, noise_domain=2)
    assert w1.get('spectral') is True and 'tx' not in w1
    w2 = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    auto = eng.simo_ber(w2, long_, rows, seed=1, fused=True, noise_domain=2).clone()
    assert 'spectral' not in w2 and 'tx' in w2
    w3 = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    forced = eng.simo_ber(w3, long_, rows, seed=1, fused=True, spectral=True, noise_domain=2).clone()
    assert w3.get('spectral') is True
    assert int((auto - forced).abs().max()) <= 3 and int(auto.sum()) > 0
