"""Fused channel + RX-FFT kernel for T > 1 transmit antennas (lte_channel_rx_fft_mimo) and the lazy-AWGN SFBC / SM
consumers, against the staged kernels (lte_channel_tdl + lte_rx_fft), which the oracle tests pin."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _engine(bw, mod, T):
    from config import LTEConfig
    from lte_b200 import LinkEngine, tables
    cfg = LTEConfig(bw, 15.0, mod)
    eng0 = LinkEngine.from_config(cfg)
    return cfg, LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(T, eng0.Np))


@pytest.mark.parametrize('bw,mod,prof,v,T,R,S,B', [
    (1.25, '16-QAM', 'Pedestrian_A', 3.0, 2, 2, 15, 3),
    (5.0, '64-QAM', 'Pedestrian_B', 3.0, 4, 3, 3, 2),         # odd antenna count: padded antenna pair
    (20.0, '64-QAM', 'Pedestrian_A', 3.0, 4, 4, 2, 2),
    (10.0, 'QPSK', 'Vehicular_A', 30.0, 2, 1, 14, 2),         # degree-2 polynomial
    (2.5, '16-QAM', 'Vehicular_A', 120.0, 2, 4, 4, 2),        # degree-4 polynomial
    (5.0, 'QPSK', 'Pedestrian_A', 3.0, 4, 8, 2, 1),           # eight receive antennas
    (5.0, 'QPSK', 'Pedestrian_A', 3.0, 8, 2, 2, 1),           # eight transmit antennas
    (1.25, '16-QAM', 'Vehicular_A', 350.0, 2, 2, 3, 2),       # degree-6 polynomial, 16 symbols per CTA
])
def test_fused_mimo_grid_and_power_match_the_staged_kernels(bw, mod, prof, v, T, R, S, B):
    from lte_b200 import chan_for
    from lte_b200 import _native as nat
    cfg, eng = _engine(bw, mod, T)
    chan = chan_for('rayleigh_mp', cfg.fs, prof, 2.0, v, gain_conversions=3 if T > 2 else 1)
    g = torch.Generator(device='cuda').manual_seed(5)
    data = torch.view_as_complex(torch.randn((B * T, S * eng.Nd, 2), generator=g, device='cuda'))
    tx, _, _ = eng.modulate(S, symbols=data, T=T, want_stats=False)
    ph = eng.random_phases(B, R * T * chan.num_taps * nat.LTE_JAKES_TONES, 3, 17)
    rx, p_ref = eng.channel(tx, chan, B, R, T=T, phases=ph)
    for window in (nat.WINDOW_USEFUL, nat.WINDOW_FULL):
        Y_ref = eng.rx_fft(rx.view(B * R, -1), B * R, S, window)
        got = eng.channel_rx_fft(tx, chan, B, R, S, ph, window, T=T)
        assert got is not None
        Y, p = got
        # at 650 Hz Doppler the STAGED kernel is the less accurate side: it linearises h over 8 samples per thread,
        # (2 pi fD / fs 3.5)^2 / 2 = 3e-5 at 1.92 MHz (channel.cu); the fused degree-6 polynomial is checked against
        # the oracle itself in tests/test_gpu_fused.py
        tol = TOL if v < 300.0 else 1e-4
        err = float((Y - Y_ref).abs().max() / Y_ref.abs().max())
        assert err < tol, err
        assert float(((p - p_ref) / p_ref).abs().max()) < tol


def test_sfbc_and_sm_passes_with_fused_link_track_the_staged_ones():
    """Same Philox draws on both paths; the only difference is the rounding of the measured stream power (the
    noise sigma) and of the grid, so the per-stream counts agree except for symbols within ~1e-6 of a boundary."""
    from core.codebook_lte import LTECodebook
    from lte_b200 import chan_for
    cfg, eng = _engine(5.0, '16-QAM', 2)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    B, R, S = 24, 2, 14
    rows = torch.tensor([10 ** (s / 10) for s in (4.0, 10.0, 16.0)], dtype=torch.float32, device='cuda').repeat(B // 3)
    rows = rows.repeat_interleave(R).contiguous()
    a = eng.sfbc_ber(chan, rows, S, R, seed=4, stream_id0=9, fused=False)
    b = eng.sfbc_ber(chan, rows, S, R, seed=4, stream_id0=9, fused=True)
    assert int(a.sum()) > 1000 and int((a - b).abs().sum()) <= 4 + int(a.sum()) // 2000
    cfg, eng = _engine(2.5, '16-QAM', 4)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
    W = LTECodebook(4, transmission_mode='TM4', rank=2).get_precoder(0)
    for det in ('MMSE', 'SIC'):
        a = eng.sm_ber(chan, W, [6.0, 14.0, 22.0] * 4, 12, 3, 4, det, seed=2, stream_id0=1, fused=False)
        b = eng.sm_ber(chan, W, [6.0, 14.0, 22.0] * 4, 12, 3, 4, det, seed=2, stream_id0=1, fused=True)
        assert int(a.sum()) > 500 and int((a - b).abs().sum()) <= 4 + int(a.sum()) // 2000, det


def test_fused_mimo_reports_unsupported_like_the_single_tx_kernel():
    from lte_b200 import chan_for
    from lte_b200 import _native as nat
    cfg, eng = _engine(1.25, 'QPSK', 2)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Vehicular_A', 2.0, 650.0)      # Doppler beyond one polynomial per symbol
    tx = torch.zeros((2, 2, 3 * eng.L), dtype=torch.complex64, device='cuda')
    ph = eng.random_phases(2, 2 * 2 * chan.num_taps * nat.LTE_JAKES_TONES, 1, 0)
    assert eng.channel_rx_fft(tx, chan, 2, 2, 3, ph, nat.WINDOW_USEFUL, T=2) is None


def test_detector_interpolating_from_pilot_estimates_equals_estimating_itself():
    """lte_crs_ls_pilots + lte_mimo_detect(Hpilot) against lte_mimo_detect estimating from Y's pilot bins, with and
    without lazy AWGN: the pilots' LS values (and noise samples) are the same numbers, so the symbols are bit-identical."""
    from core.codebook_lte import LTECodebook
    from lte_b200 import _native as nat
    for bw, mod, T, rank, det in ((1.25, '16-QAM', 4, 2, 'MMSE'), (5.0, '64-QAM', 2, 2, 'ZF'), (2.5, 'QPSK', 4, 3, 'SIC')):
        cfg, eng = _engine(bw, mod, T)
        W = LTECodebook(T, transmission_mode='TM4', rank=rank).get_precoder(0)
        B, R, S = 3, 4, 2
        k0, nk = eng.window(nat.WINDOW_USEFUL)
        g = torch.Generator(device='cuda').manual_seed(2)
        Y = torch.view_as_complex(torch.randn((B * R, S, nk, 2), generator=g, device='cuda'))
        power = torch.full((B, R), 3.0e3, dtype=torch.float64, device='cuda')
        snr = torch.full((B * R,), 20.0, dtype=torch.float32, device='cuda')
        for awgn in (None, eng.awgn_desc(power, snr, 5, 40)):
            Hp = eng.estimate_pilots(Y, B * R, S, nat.WINDOW_USEFUL, awgn=awgn)
            a = eng.mimo_detect(Y, None, W, 0.05, det, B, R, S, nat.WINDOW_USEFUL, awgn=awgn)
            b = eng.mimo_detect(Y, None, W, 0.05, det, B, R, S, nat.WINDOW_USEFUL, awgn=awgn, Hpilot=Hp)
            assert torch.equal(torch.view_as_real(a), torch.view_as_real(b)), (bw, det, awgn is not None)
