"""Beamforming path on the GPU: lte_bf_weights / lte_bf_link vs the oracle and the
golden vectors from the reference's OFDMSimulator.simulate_beamforming; reference-shaped helper
classes; the sharded sweep."""
import numpy as np
import pytest
import torch

from cases import BF_CASES
from helpers import golden_bits, golden_bits_rx, load_golden, numerology, rel_err
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('case', BF_CASES, ids=lambda c: c['name'])
def test_simulate_beamforming_matches_reference(case):
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    from gpu_chain import boundary_distance
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    sim = OFDMSimulator(LTEConfig(case['bw'], 15.0, case['mod']), channel_type='awgn')
    for snr in case['snrs']:
        np.random.seed(case['gseed'])
        r = sim.simulate_beamforming(bits, snr_db=snr, num_tx=case['T'], num_rx=case['R'], codebook_type=case['cb'],
                                     velocity_kmh=case['v'], update_mode=case['upd'])
        o = O.simulate_beamforming(bits, snr, num, case['T'], case['R'], case['upd'], global_seed=case['gseed'])
        assert rel_err(r['channel_matrix'], g[f'channel_matrix_{snr}']) < 1e-12          # same draws
        assert r['pmi_history'] == list(g[f'pmi_history_{snr}'])
        assert abs(r['beamforming_gain_db'] - g[f'gain_unique_{snr}'][0]) < 1e-4
        assert r['unique_pmis'] == int(g[f'gain_unique_{snr}'][1])
        assert rel_err(r['precoder'], o['W']) < 1e-6
        n = len(r['symbols_rx'])
        assert rel_err(r['symbols_rx'], o['symbols'][:n]) < TOL                           # equalised symbols, fp32
        want = golden_bits_rx(g, snr)
        diff = np.flatnonzero(r['bits_received_array'] != want)
        if len(diff):             # only decisions whose fp64 value sits on a slicer boundary may differ
            d = boundary_distance(o['symbols'], num.modulation)
            bad = np.unique(diff // num.bits_per_symbol)
            assert np.all(d[bad] < 1e-5 * np.maximum(1.0, np.abs(o['symbols'][bad])))
        assert abs(r['errors'] - int(g[f'errors_{snr}'])) <= len(diff)
        assert r['errors'] == int(np.sum(r['bits_received_array'] != bits))
        assert (r['mode'], r['num_tx'], r['num_rx'], r['codebook_type']) == ('Beamforming', case['T'], case['R'],
                                                                              case['cb'])
        assert r['transmitted_bits'] == len(bits) == r['received_bits']


@pytest.mark.parametrize('T,R', [(2, 1), (4, 2), (8, 4), (8, 8), (4, 1)])
def test_bf_weights_kernel_matches_oracle(T, R):
    from core.csi_feedback import _bf_engine
    eng = _bf_engine()
    rs = np.random.RandomState(10 * T + R)
    B = 257
    H = ((rs.randn(B, R, T) + 1j * rs.randn(B, R, T)) / np.sqrt(2)).astype(np.complex64)
    cb = O.codebook(T, 1)
    h = torch.from_numpy(H).cuda()
    for mode in ('MRT', 'CODEBOOK'):
        W, heff, pmi, gain = eng.bf_weights(h, cb, mode=mode)
        Wn, hn, pn, gn = W.cpu().numpy(), heff.cpu().numpy(), pmi.cpu().numpy(), gain.cpu().numpy()
        for b in range(B):
            Hb = H[b].astype(complex)
            best, _ = O.select_best_pmi(Hb, cb)
            assert pn[b] == best
            Wo = O.mrt_weights(Hb) if mode == 'MRT' else cb[best]
            assert rel_err(Wn[b], Wo) < 2e-7
            assert rel_err(hn[b], Hb @ Wo) < 2e-7
            assert abs(gn[b] - O.beamforming_gain_db(Hb, Wo, T)) < 1e-4
    if R == 1:                                    # MRT on a single RX antenna: array gain is exactly 10 log10(T)
        _, _, _, gain = eng.bf_weights(h, cb, mode='MRT')
        assert np.allclose(gain.cpu().numpy(), 10 * np.log10(T), atol=1e-4)


def test_helper_classes_follow_the_reference():
    from core.beamforming_precoder import AdaptiveBeamforming, BeamformingPrecoder
    from core.csi_feedback import CSIFeedback
    rs = np.random.RandomState(3)
    H = (rs.randn(2, 4) + 1j * rs.randn(2, 4)) / np.sqrt(2)
    csi = CSIFeedback(4, 2, codebook_type='TM6')
    fb = csi.generate_feedback(H, noise_variance=0.5)
    best, metric = O.select_best_pmi(H.astype(np.complex64).astype(complex), O.codebook(4, 1))
    assert fb['pmi'] == best and np.array_equal(fb['precoder'], O.codebook(4, 1)[best])
    assert fb['cqi'] == O.sinr_to_cqi(fb['sinr_db']) and abs(fb['sinr_db'] - 10 * np.log10(metric / 0.5)) < 1e-5
    assert fb['ri'] in (1, 2) and csi.get_statistics()['total_feedbacks'] == 1
    p = BeamformingPrecoder(4)
    assert p.calculate_beamforming_gain(H) == 0.0
    with pytest.raises(ValueError):
        p.apply_precoding(np.ones(4, complex))
    W = p.update_precoder(H, method='MRT')
    assert W.shape == (4, 1) and rel_err(W, O.mrt_weights(H)) < 2e-7
    assert abs(p.calculate_beamforming_gain(H) - O.beamforming_gain_db(H, O.mrt_weights(H), 4)) < 1e-5
    s = rs.randn(3000) + 1j * rs.randn(3000)                     # longer than one plan row
    assert rel_err(p.apply_precoding(s), W @ s.reshape(1, -1)) < 2e-7
    assert rel_err(p.apply_precoding(s[:37], O.codebook(4, 1)[5]), O.codebook(4, 1)[5] @ s[:37].reshape(1, -1)) < 2e-7
    We = p.update_precoder(H, method='eigen')
    ev = np.linalg.eigvalsh(H.conj().T @ H)[-1]
    assert abs(np.sum(np.abs(H @ We) ** 2) - ev) < 1e-9
    with pytest.raises(ValueError):
        p.update_precoder(H, method='nope')
    a = AdaptiveBeamforming(2, 3.0, 2.0)
    assert a.update_period == O.beamforming_update_period(3.0) == 48
    assert AdaptiveBeamforming(2, 0.0, 2.0).update_period == 100
    x = a.process_symbol(s[:10], H[:, :2])
    assert x.shape == (2, 10) and a.symbols_since_update == 1 and not a.should_update()


def test_sweep_is_invariant_to_batching_and_sharding():
    from config import LTEConfig
    from lte_b200 import LinkEngine
    from lte_b200.sweep import beamforming_sweep
    eng = LinkEngine.from_config(LTEConfig(1.25, 15.0, '16-QAM'))
    cb = O.codebook(4, 1)
    snrs = [0.0, 6.0, 12.0]
    one = beamforming_sweep(eng, cb, snrs, 40, 4, 2, mode='CODEBOOK', symbols_per_stream=3, seed=5, batch_trials=40)
    parts = [beamforming_sweep(eng, cb, snrs, 40, 4, 2, mode='CODEBOOK', symbols_per_stream=3, seed=5, batch_trials=7,
                               rank=r, world=3) for r in range(3)]
    assert torch.equal(sum(p['errors'] for p in parts), one['errors'])
    assert torch.equal(sum(p['pmi_hist'] for p in parts), one['pmi_hist'])
    assert int(one['pmi_hist'].sum()) == 40 * len(snrs)
    assert one['ber'][0] > one['ber'][1] > one['ber'][2] >= 0
    # MRT beats the quantised codebook; a single RX antenna gets the full 10 log10(T) array gain
    mrt = beamforming_sweep(eng, cb, snrs, 40, 4, 1, mode='MRT', symbols_per_stream=3, seed=5)
    cbk = beamforming_sweep(eng, cb, snrs, 40, 4, 1, mode='CODEBOOK', symbols_per_stream=3, seed=5)
    assert abs(mrt['mean_gain_db'] - 10 * np.log10(4)) < 1e-3 and cbk['mean_gain_db'] < mrt['mean_gain_db']
    assert int(mrt['errors'].sum()) <= int(cbk['errors'].sum())


def test_philox_mode_gives_independent_trials_and_plausible_ber():
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    sim = OFDMSimulator(LTEConfig(1.25, 15.0, 'QPSK'), channel_type='awgn', rng='philox', seed=9)
    bits = np.random.RandomState(1).randint(0, 2, 62 * 2 * 4)
    a = sim.simulate_beamforming(bits, snr_db=5.0, num_tx=2, num_rx=2)
    b = sim.simulate_beamforming(bits, snr_db=5.0, num_tx=2, num_rx=2)
    assert not np.allclose(a['channel_matrix'], b['channel_matrix'])
    hi = sim.simulate_beamforming(bits, snr_db=40.0, num_tx=8, num_rx=1)
    assert hi['errors'] == 0 and abs(hi['beamforming_gain_db'] - 10 * np.log10(8)) < 1e-3
    with pytest.raises(ValueError):
        sim.simulate_beamforming([], snr_db=5.0)
