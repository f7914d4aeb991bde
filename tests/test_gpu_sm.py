"""Spatial multiplexing (BASELINE.json config 5) on the GPU: detector kernels vs the oracle's
numpy.linalg restatement, API vs golden vectors from the reference."""
import numpy as np
import pytest

from cases import SM_CASES
from helpers import golden_bits, golden_bits_rx, load_golden, numerology, rel_err
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('det,L,R,T', [('MMSE', 4, 4, 4), ('MMSE', 2, 4, 4), ('ZF', 2, 4, 4), ('ZF', 4, 4, 4),
                                       ('SIC', 3, 4, 4), ('SIC', 2, 2, 2), ('MRC', 1, 2, 4), ('MMSE', 4, 8, 8)])
def test_detector_kernel_matches_numpy(det, L, R, T):
    from core.mimo_detector import MIMODetector
    rs = np.random.RandomState(R * 100 + L)
    K = 300
    H = (rs.randn(R, T, K) + 1j * rs.randn(R, T, K)) / np.sqrt(2)
    W = O.codebook(T, L)[min(1, len(O.codebook(T, L)) - 1)]
    c = O.constellation('16-QAM')
    s = c[rs.randint(0, 16, (L, K))]
    y = np.einsum('rtk,tl,lk->rk', H, W, s) + 0.05 * (rs.randn(R, K) + 1j * rs.randn(R, K))
    # the kernel sees fp32 inputs: give the oracle the same rounded values
    H32, y32 = H.astype(np.complex64).astype(complex), y.astype(np.complex64).astype(complex)
    W32 = W.astype(np.complex64).astype(complex)
    want = np.stack([O.mimo_detect(y32[:, k], H32[:, :, k] @ W32, 0.01, det, c) for k in range(K)], axis=1)
    d = MIMODetector(R, L, det, constellation=c)
    got = d.detect(y, H, 0.01, W)
    assert got.shape == (L, K)
    if det == 'SIC':
        assert np.mean(np.abs(got - want) < 1e-6) > 0.995          # hard decisions: identical but for boundary cases
    else:
        assert rel_err(got, want) < TOL
    one = d.detect(y[:, 0], H[:, :, 0], 0.01, W)
    assert one.shape == (L,) and np.allclose(one, got[:, 0], atol=1e-5)


def test_detector_error_conventions():
    from core.mimo_detector import MIMODetector
    with pytest.raises(ValueError):
        MIMODetector(2, 4, 'MMSE')
    d = MIMODetector(4, 2, 'MRC')
    with pytest.raises(ValueError):
        d.detect(np.zeros((4, 3), complex), np.zeros((4, 4, 3), complex), 0.1, np.zeros((4, 2), complex))
    with pytest.raises(ValueError):
        MIMODetector(4, 2, 'ML').detect(np.zeros((4, 3), complex), np.zeros((4, 4, 3), complex), 0.1)


def test_reference_toy_case_2x2_mmse():
    """reference core/mimo_detector.py:387-404."""
    from core.mimo_detector import MIMODetector
    H = np.array([[1.0 + 0.5j, 0.3 - 0.2j], [0.2 + 0.1j, 0.9 - 0.3j]])
    s = np.array([1 + 1j, -1 + 1j]) / np.sqrt(2)
    y = H @ s + 0.1 * (np.array([0.3, -0.2]) + 1j * np.array([0.1, 0.25]))
    for det in ('MMSE', 'ZF'):
        assert np.linalg.norm(s - MIMODetector(2, 2, det).detect(y, H, 0.01)) < 1.0


@pytest.mark.parametrize('case', SM_CASES, ids=lambda c: c['name'])
def test_simulate_spatial_multiplexing_matches_reference(case):
    from config import LTEConfig
    from core.ofdm_core import simulate_spatial_multiplexing
    from gpu_chain import boundary_distance
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    cfg = LTEConfig(case['bw'], 15.0, case['mod'])
    for snr in case['snrs']:
        np.random.seed(case['gseed'])
        r = simulate_spatial_multiplexing(bits, num_tx=case['T'], num_rx=case['R'], rank=case['rank'],
                                          detector_type=case['det'], modulation=case['mod'], snr_db=snr, config=cfg,
                                          channel_type=case['ch'], itu_profile=case['prof'],
                                          velocity_kmh=case['v'], frequency_ghz=2.0)
        o = O.simulate_sm(bits, num, case['T'], case['R'], case['rank'], case['det'], snr, case['ch'], case['prof'],
                          case['v'], 2.0, global_seed=case['gseed'])
        assert [r['rank'], r['pmi_used']] == list(g[f'rank_pmi_{snr}'])
        assert rel_err(r['precoder_matrix'], g[f'W_{snr}']) < 1e-12
        assert rel_err(r['channel_matrix'], g[f'channel_matrix_{snr}']) < TOL
        want = golden_bits_rx(g, snr)
        diff = np.flatnonzero(r['bits_received_array'] != want)
        nsym_used = len(o['symbols'])
        if case['det'] != 'SIC':
            # detector outputs: median element-wise agreement; ill-conditioned bins amplify fp32 input rounding
            elem = np.abs(r['symbols_rx'][:nsym_used] - o['symbols']) / np.maximum(np.abs(o['symbols']), 1e-3)
            assert np.median(elem) < TOL
            if len(diff):
                d = boundary_distance(o['symbols'], num.modulation)
                bad = np.unique(diff // num.bits_per_symbol)
                assert np.all(d[bad] < 1e-3 * np.maximum(1.0, np.abs(o['symbols'][bad])))
            assert len(diff) <= 3 * num.bits_per_symbol
        else:
            assert len(diff) <= 0.002 * len(want) + 2 * num.bits_per_symbol    # a flipped hard decision propagates
        assert abs(r['errors'] - int(g[f'errors_{snr}'])) <= len(diff)
        assert r['mode'] == 'Spatial Multiplexing TM4' and r['num_tx'] == case['T'] and r['detector_type'] == case['det']


def test_helper_classes():
    from core.codebook_lte import LTECodebook
    from core.layer_mapper import LayerMapper
    from core.rank_adaptation import RankAdaptation
    for T in (2, 4, 8):
        for r in range(1, min(T, 4) + 1):
            cb = LTECodebook(T, 'TM4', r)
            ref = O.codebook(T, r)
            assert cb.codebook_size == len(ref)
            for a, b in zip(cb.get_codebook(), ref):
                assert np.array_equal(np.asarray(a, dtype=complex), np.asarray(b, dtype=complex))
    with pytest.raises(ValueError):
        LTECodebook(4, 'TM6', 2)
    with pytest.raises(ValueError):
        LTECodebook(2, 'TM4', 3)
    with pytest.raises(ValueError):
        LTECodebook(4, 'TM4', 1).get_precoder(99)
    m = LayerMapper(3)
    x = np.arange(10) + 0j
    assert np.array_equal(m.demap_from_layers(m.map_to_layers(x), 10), x)
    assert m.get_padded_length(10) == 12 and m.get_symbols_per_layer(10) == 4
    with pytest.raises(ValueError):
        LayerMapper(9)
    H = (np.random.RandomState(0).randn(4, 4) + 1j * np.random.RandomState(1).randn(4, 4)) / np.sqrt(8)
    fb = RankAdaptation(4, 4, snr_db=20.0).get_feedback(H)
    ri, pmi, W = O.rank_feedback(H, 4, 4, 20.0)
    assert (fb['ri'], fb['pmi']) == (ri, pmi) and np.array_equal(fb['W'], W)
    assert RankAdaptation(4, 4, snr_db=3.0).get_feedback(H)['ri'] == 1


def test_sm_sweep_fixed_and_adaptive_rank():
    """BASELINE config 5 as a batched sweep: sharding invariance (fixed and adaptive rank), the rank
    histogram of the adaptive rule, and statistical agreement with looping simulate_spatial_multiplexing."""
    import torch
    from config import LTEConfig
    from core.codebook_lte import LTECodebook
    from core.ofdm_core import simulate_spatial_multiplexing
    from core.rank_adaptation import RankAdaptation
    from lte_b200.sweep import sm_sweep
    cfg = LTEConfig(1.25, 15.0, '16-QAM')
    snrs = [8.0, 20.0]

    def precoder(r):
        return LTECodebook(4, transmission_mode='TM4', rank=r).get_precoder(0)

    def feedback(H, snr_db):
        fb = RankAdaptation(4, 4, snr_db=snr_db).get_feedback(H)
        return fb['ri'], fb['pmi'], fb['W']

    kw = dict(num_tx=4, num_rx=4, detector='MMSE', symbols_per_stream=2, seed=5)
    one = sm_sweep(cfg, snrs, 40, rank=2, precoder=precoder, batch_trials=40, **kw)
    parts = [sm_sweep(cfg, snrs, 40, rank=2, precoder=precoder, batch_trials=9, rank_id=r, world=2, **kw) for r in range(2)]
    assert torch.equal(parts[0]['errors'] + parts[1]['errors'], one['errors'])
    assert one['rank_hist'][:, 1].tolist() == [40, 40] and int(one['rank_hist'].sum()) == 80
    assert one['ber'][0] > one['ber'][1]
    ad = sm_sweep(cfg, snrs, 40, rank='adaptive', feedback=feedback, feedback_block=8, batch_trials=40, **kw)
    ad2 = [sm_sweep(cfg, snrs, 40, rank='adaptive', feedback=feedback, feedback_block=8, batch_trials=7, rank_id=r,
                    world=3, **kw) for r in range(3)]
    assert torch.equal(sum(p['errors'] for p in ad2), ad['errors'])
    assert torch.equal(sum(p['rank_hist'] for p in ad2), ad['rank_hist'])
    assert ad['rank_hist'].sum(dim=1).tolist() == [40, 40]
    # per-call API with independent Philox draws: same BER regime at rank 2, 20 dB
    bits = np.random.RandomState(0).randint(0, 2, 2 * 62 * 4)
    loop = np.mean([simulate_spatial_multiplexing(bits, num_tx=4, num_rx=4, rank=2, detector_type='MMSE',
                                                  modulation='16-QAM', snr_db=20.0, config=cfg,
                                                  channel_type='rayleigh_mp', rng='philox', seed=i)['ber']
                    for i in range(120)])
    big = sm_sweep(cfg, [20.0], 240, rank=2, precoder=precoder, **kw)['ber'].numpy()[0]
    assert abs(loop - big) / max(big, 1e-9) < 0.35
    # the same with the feedback rule on the device (the default): sharding invariance and the SNR gate
    dv = sm_sweep(cfg, [3.0, 8.0, 20.0], 40, rank='adaptive', feedback_block=8, batch_trials=40, **kw)
    dv2 = [sm_sweep(cfg, [3.0, 8.0, 20.0], 40, rank='adaptive', feedback_block=8, batch_trials=7, rank_id=r, world=3, **kw)
           for r in range(3)]
    assert torch.equal(sum(p['errors'] for p in dv2), dv['errors'])
    assert torch.equal(sum(p['rank_hist'] for p in dv2), dv['rank_hist'])
    assert dv['rank_hist'].sum(dim=1).tolist() == [40, 40, 40]
    assert dv['rank_hist'][0].tolist() == [40, 0, 0, 0]              # below 5 dB: rank 1
    assert int(dv['rank_hist'][1, 2:].sum()) == 0                   # below 10 dB: at most rank 2
    with pytest.raises(ValueError):
        sm_sweep(cfg, snrs, 4, rank=2, **kw)                        # a fixed rank needs its precoder


@pytest.mark.parametrize('R,T', [(4, 4), (2, 4), (8, 4), (2, 2), (4, 2)])
def test_rank_feedback_kernel_matches_the_reference_rule(R, T):
    """lte_rank_feedback against RankAdaptation.get_feedback (core/rank_adaptation.py, the reference's rule) on the
    same matrices: RI and PMI equal for every matrix and SNR point, including the SNR gates and weak channels."""
    import torch
    from core.rank_adaptation import RankAdaptation
    from lte_b200 import LinkEngine
    from config import LTEConfig
    eng = LinkEngine.from_config(LTEConfig(1.25, 15.0, 'QPSK'))
    n = 600
    H = eng.random_channel(n, R, T, 11) * (1.0 / np.sqrt(T))
    Hn = H.cpu().numpy().astype(np.complex128)
    Hn[:5] *= 1e-7                                                  # lambda_max < 1e-10: rank 1
    H = torch.from_numpy(Hn.astype(np.complex64)).to(eng.device)
    Hn = H.cpu().numpy().astype(np.complex128)
    snr = np.random.RandomState(3).choice([-3.0, 4.9, 5.0, 7.5, 10.0, 15.0, 22.0, 30.0], n)
    max_rank = min(R, T, 4)
    tab, sizes = eng.rank_codebook(T, max_rank)
    ri, pmi = eng.rank_feedback(H, torch.from_numpy(snr).to(eng.device), tab, sizes, max_rank=max_rank)
    ri, pmi = ri.cpu().numpy(), pmi.cpu().numpy()
    from core.codebook_lte import LTECodebook
    seen, decided = set(), 0
    for i in range(n):
        fb = RankAdaptation(T, R, snr_db=float(snr[i])).get_feedback(Hn[i])
        assert int(ri[i]) == int(fb['ri']), (i, snr[i])
        # PMI: the same entry wherever the capacity metric has a winner.  A 1e-7 channel, or rank T with its unitary
        # precoders (det(I + c H W W^H H^H) does not depend on W), leave only rounding noise between the entries:
        # there the kernel's choice must be one of the (numerically) tied maxima.
        cb = LTECodebook(T, transmission_mode='TM4', rank=int(ri[i]))
        c = 10 ** (float(snr[i]) / 10) / int(ri[i])
        m = np.array([np.log2(np.linalg.det(np.eye(R) + c * (Hn[i] @ cb.get_precoder(k)) @ (Hn[i] @ cb.get_precoder(k)).conj().T).real)
                      for k in range(cb.codebook_size)])
        tied = np.flatnonzero(m >= m.max() - 1e-9 * max(1.0, abs(m.max())))
        assert int(pmi[i]) in tied, (i, snr[i], m)
        if len(tied) == 1:
            assert int(pmi[i]) == int(fb['pmi']), (i, snr[i])
            decided += 1
        seen.add(int(ri[i]))
    assert seen == set(range(1, max_rank + 1))
    assert decided > n // 3


def test_sm_ber_per_stream_snr_equals_per_snr_passes():
    """lte_mimo_detect with one sigma^2 per stream: a pass whose streams run at different SNR points gives, stream
    by stream, the counts of separate single-SNR passes on the same stream ids (so a sweep may put all its SNR
    points into one launch)."""
    import torch
    from config import LTEConfig
    from core.codebook_lte import LTECodebook
    from lte_b200 import LinkEngine, chan_for, tables
    cfg = LTEConfig(1.25, 15.0, '16-QAM')
    eng0 = LinkEngine.from_config(cfg)
    eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(4, eng0.Np))
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
    W = LTECodebook(4, transmission_mode='TM4', rank=2).get_precoder(0)
    snrs = [4.0, 12.0, 25.0]
    for det in ('MMSE', 'SIC'):
        B = 6
        per_stream = [snrs[b % 3] for b in range(B)]
        batched = eng.sm_ber(chan, W, per_stream, B, 2, 4, det, seed=3, stream_id0=50).cpu()
        for b in range(B):
            single = eng.sm_ber(chan, W, per_stream[b], 1, 2, 4, det, seed=3, stream_id0=50 + b).cpu()
            assert int(single[0]) == int(batched[b]), (det, b)
    with pytest.raises(ValueError):
        eng.sm_ber(chan, W, [4.0, 5.0], 3, 2, 4, 'MMSE', seed=3)


def test_detector_estimating_from_the_pilot_bins_equals_the_estimate_tensor():
    """lte_mimo_detect(H = NULL) forms the per-symbol CRS estimates from Y's pilot bins itself: symbols and counts
    are bit-identical to the explicit lte_crs_ls_interp passes + H tensor, for every detector."""
    import torch
    from config import LTEConfig
    from core.codebook_lte import LTECodebook
    from lte_b200 import LinkEngine, chan_for, tables
    from lte_b200 import _native as nat
    for bw, mod, T, rank in ((1.25, '16-QAM', 4, 2), (5.0, '64-QAM', 2, 2), (2.5, 'QPSK', 4, 4)):
        cfg = LTEConfig(bw, 15.0, mod)
        eng0 = LinkEngine.from_config(cfg)
        eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(T, eng0.Np))
        chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
        W = LTECodebook(T, transmission_mode='TM4', rank=rank).get_precoder(0)
        for det in ('MMSE', 'ZF', 'SIC'):
            a = eng.sm_ber(chan, W, 14.0, 5, 3, 4, det, seed=9, stream_id0=2, estimate_tensor=True)
            b = eng.sm_ber(chan, W, 14.0, 5, 3, 4, det, seed=9, stream_id0=2)
            assert torch.equal(a, b), (bw, det)
        # symbols, not only counts
        B, R, S = 3, 4, 2
        k0, nk = eng.window(nat.WINDOW_USEFUL)
        g = torch.Generator(device='cuda').manual_seed(1)
        Y = torch.view_as_complex(torch.randn((B * R, S, nk, 2), generator=g, device='cuda'))
        H = torch.empty((T, B * R, S, nk), dtype=torch.complex64, device='cuda')
        for t in range(T):
            eng.estimate(Y.view(B * R * S, 1, nk), B * R * S, 1, nat.WINDOW_USEFUL, pilot_set=t, out=H[t].view(B * R * S, 1, nk))
        s1 = eng.mimo_detect(Y, H, W, 0.05, 'MMSE', B, R, S, nat.WINDOW_USEFUL)
        s2 = eng.mimo_detect(Y, None, W, 0.05, 'MMSE', B, R, S, nat.WINDOW_USEFUL)
        assert torch.equal(torch.view_as_real(s1), torch.view_as_real(s2))
