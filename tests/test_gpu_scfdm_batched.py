"""Batched SC-FDM BER chain (BASELINE config 2, LinkEngine.siso_ber(sc_fdm=True) / sweep.scfdm_sweep): per-stream
error counts must equal the per-call API (reference simulate_siso with enable_sc_fdm=True,
core/ofdm_core.py:660-737, core/dft_precoding.py:67-93, core/lte_receiver.py:319-333) on the same draws."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('bw,mod,prof,v', [(10.0, '16-QAM', 'Pedestrian_A', 3.0), (5.0, 'QPSK', 'Vehicular_A', 30.0)])
def test_batched_scfdm_counts_equal_the_per_call_api(bw, mod, prof, v):
    from config import LTEConfig
    from core import _backend as be
    from core.ofdm_core import OFDMSimulator
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(bw, 15.0, mod)
    sim = OFDMSimulator(cfg, channel_type='rayleigh_mp', mode='lte', enable_sc_fdm=True, itu_profile=prof,
                        frequency_ghz=2.0, velocity_kmh=v, rng='philox', seed=21)
    eng = LinkEngine.from_config(cfg)
    S = 3
    nbits = S * eng.Nd * eng.bps - 7
    bits = np.random.RandomState(5).randint(0, 2, nbits)
    snrs = [6.0, 14.0, 22.0, 30.0]
    sid0 = next(be._PHILOX_CALLS) + 1                     # the next per-call stream id
    want = [sim.simulate_siso(bits, snr_db=s)['bit_errors'] for s in snrs]
    chan = chan_for('rayleigh_mp', cfg.fs, prof, 2.0, v)
    idx = eng.bits_to_indices(torch.from_numpy(bits.astype(np.uint8)).cuda()[None], nbits, S).expand(len(snrs), -1).contiguous()
    rows = torch.tensor([10 ** (s / 10) for s in snrs], dtype=torch.float32, device='cuda')
    got = eng.siso_ber(chan, rows, S, 21, stream_id0=sid0, idx=idx, nbits=nbits, sc_fdm=True, noise_domain=0)
    assert [int(x) for x in got.cpu()] == want
    assert sum(want) > 0


def test_scfdm_sweep_ber_and_papr_in_one_pass():
    """BER falls with SNR, the PAPR histogram counts every OFDM symbol once, SC-FDM's PAPR sits below OFDM's
    (by ~0.75 dB with the reference's grid -- DFT-spread symbols interleaved with CRS pilots around a nulled DC
    bin -- not the 3-4 dB its README quotes), and the sweep does not depend on the batch size."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    from lte_b200.sweep import scfdm_sweep
    cfg = LTEConfig(10.0, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    snr = [0.0, 10.0, 20.0, 30.0]
    a = scfdm_sweep(eng, chan, snr, 24, seed=3, batch_trials=24)
    b = scfdm_sweep(eng, chan, snr, 24, seed=3, batch_trials=5)
    assert torch.equal(a['errors'], b['errors']) and torch.equal(a['papr_hist'], b['papr_hist'])
    assert int(a['papr_hist'].sum()) == 24 * len(snr) * 14
    assert a['ber'][0] > a['ber'][-1]
    o = scfdm_sweep(eng, chan, snr, 24, seed=3, batch_trials=24, sc_fdm=False)
    centre = a['papr_edges_db'][:-1] + 0.05
    mean_sc = float((a['papr_hist'].double() * centre).sum() / a['papr_hist'].sum())
    mean_of = float((o['papr_hist'].double() * centre).sum() / o['papr_hist'].sum())
    assert mean_sc < mean_of - 0.3


def test_run_ber_sweep_routes_scfdm_through_the_batched_chain():
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    from lte_b200 import LinkEngine
    cfg = LTEConfig(5.0, 15.0, 'QPSK')
    sim = OFDMSimulator(cfg, channel_type='rayleigh_mp', mode='lte', enable_sc_fdm=True, itu_profile='Pedestrian_A',
                        frequency_ghz=2.0, velocity_kmh=3.0, rng='philox', seed=4)
    eng = sim._engine()
    l0 = eng.launches
    np.random.seed(0)
    r = sim.run_ber_sweep(4000, [0.0, 10.0, 20.0, 30.0], num_trials=8)
    assert eng.launches - l0 < 40                       # one batch of 32 streams, not 32 per-call chains
    assert r['ber_mean'][0] > r['ber_mean'][-1]
    assert isinstance(eng, LinkEngine)


def test_dft_qam_equals_the_separate_calls_and_the_fused_link_tracks_the_staged_one():
    """lte_dft_qam = qam_map + dft_m (same arithmetic: bit-identical); the SISO pass through the fused channel + FFT
    kernel with lazy AWGN in the estimator and the equaliser gives the staged pass's counts up to symbols on a slicer
    boundary (the measured stream power differs in its last bits), for OFDM and SC-FDM."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    for bw, mod in ((10.0, '16-QAM'), (1.25, '64-QAM'), (5.0, 'QPSK')):
        cfg = LTEConfig(bw, 15.0, mod)
        eng = LinkEngine.from_config(cfg)
        B, S = 6, 15
        idx = eng.random_indices(B, S, 3, 0)
        a = eng.dft_m(eng.qam_map(idx).view(B * S, eng.Nd), eng.Nd)
        b = eng.dft_qam(idx.view(B * S, eng.Nd), eng.Nd)
        assert torch.equal(torch.view_as_real(a), torch.view_as_real(b))
        chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
        rows = torch.tensor([10 ** (s / 10) for s in (6.0, 16.0, 26.0)], dtype=torch.float32, device='cuda').repeat(B // 3).contiguous()
        for sc in (True, False):
            e_staged = eng.siso_ber(chan, rows, S, 9, stream_id0=4, idx=idx, sc_fdm=sc, fused=False)
            e_fused = eng.siso_ber(chan, rows, S, 9, stream_id0=4, idx=idx, sc_fdm=sc, fused=True)
            assert int(e_staged.sum()) > 0 and int((e_staged - e_fused).abs().sum()) <= 4 + int(e_staged.sum()) // 1000
