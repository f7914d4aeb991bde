"""PAPR / CCDF engine on the GPU against the reference's golden values and the oracle:
fused TX+PAPR epilogue, PAPR of an existing stream, histogram, the reference-shaped OFDMSystem and
the sharded sweep."""
import numpy as np
import pytest

from cases import PAPR_CASES
from helpers import load_golden
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu

DB_TOL = 2e-4        # dB; fp32 peak/mean ratio (1e-5 relative = 4.3e-5 dB) plus log10f rounding


def engine_for(case):
    from lte_b200 import LinkEngine
    num = O.Numerology(case['bw'], 15.0, case['mod'])
    return LinkEngine(num.N, num.Nc, num.cp_length, num.bits_per_symbol, num.fs), num


@pytest.mark.parametrize('case', PAPR_CASES, ids=[c['name'] for c in PAPR_CASES])
def test_fused_tx_papr_matches_reference(case):
    import torch
    g = load_golden(case['name'])
    eng, num = engine_for(case)
    bits = torch.from_numpy(g['bits'].astype(np.uint8)).cuda().reshape(1, -1)
    S = eng.symbols_for_bits(bits.shape[1])
    idx = eng.bits_to_indices(bits, bits.shape[1], S)
    hist = torch.zeros(300, dtype=torch.int64, device='cuda')
    db, pm, tx = eng.modulate_papr(S, idx=idx, write_tx=True, hist=hist, hist_lo=0.0, hist_step=0.05,
                                   want_peak_mean=True)
    db = db.cpu().numpy().reshape(-1).astype(np.float64)
    assert np.max(np.abs(db - g['ofdm_no_cp_db'])) < DB_TOL
    # the stream written alongside is the ordinary TX output
    tx_ref, _, _ = eng.modulate(S, idx=idx)
    assert torch.equal(tx, tx_ref)
    # peak / mean of the useful part against the oracle on the oracle's own signal
    sig, _ = O.modulate_stream(g['bits'], num)
    _, peak, mean = O.papr_per_symbol_db(sig, num)
    pmn = pm.cpu().numpy().reshape(-1, 2)
    assert np.allclose(pmn[:, 0], peak, rtol=1e-5) and np.allclose(pmn[:, 1], mean, rtol=1e-5)
    # histogram = the documented bin rule applied to the values the kernel reported
    want = O.papr_histogram(db, 0.0, 0.05, 300)
    assert np.array_equal(hist.cpu().numpy(), want)
    # sweep mode: no stream written, same values
    db2, _, none = eng.modulate_papr(S, idx=idx)
    assert none is None and np.array_equal(db2.cpu().numpy().reshape(-1).astype(np.float64), db)

    # SC-FDM: constellation map + M-point DFT + symbols path
    pre = eng.dft_m(eng.qam_map(idx).reshape(S, eng.Nd), eng.Nd)
    db3, _, _ = eng.modulate_papr(S, symbols=pre.reshape(1, -1))
    assert np.max(np.abs(db3.cpu().numpy().reshape(-1) - g['scfdm_no_cp_db'])) < DB_TOL


@pytest.mark.parametrize('case', PAPR_CASES[:2], ids=[c['name'] for c in PAPR_CASES[:2]])
def test_ofdm_system_api_matches_reference(case):
    from config import LTEConfig
    from core.ofdm_system import OFDMSystem
    g = load_golden(case['name'])
    for tag, flag in (('ofdm', False), ('scfdm', True)):
        sysm = OFDMSystem(LTEConfig(case['bw'], 15.0, case['mod']), 'awgn', mode='lte', enable_sc_fdm=flag)
        sig = sysm.modulator.modulate_stream(g['bits'])[0]
        a = sysm.calculate_papr_without_cp(sig)
        assert np.max(np.abs(a['papr_per_symbol'] - g[f'{tag}_no_cp_db'])) < DB_TOL
        assert a['num_symbols'] == len(g[f'{tag}_no_cp_db']) and len(a['papr_values']) == a['num_symbols']
        assert abs(a['papr_mean'] - g[f'{tag}_no_cp_db'].mean()) < DB_TOL
        b = sysm.calculate_papr_per_symbol(sig)
        assert np.max(np.abs(b['papr_per_symbol'] - g[f'{tag}_cp_db'])) < DB_TOL
        assert np.allclose(b['power_peak_per_symbol'], g[f'{tag}_cp_peak'], rtol=1e-5)
        assert np.allclose(b['power_avg_per_symbol'], g[f'{tag}_cp_avg'], rtol=1e-5)
        c = sysm.calculate_papr(sig)
        assert abs(c['papr_db'] - g[f'{tag}_stream'][0]) < DB_TOL
        # the whole link through the same object: noiseless enough at 60 dB to return the bits
        r = sysm.transmit(g['bits'].astype(np.int64), snr_db=60.0)
        assert r['bit_errors'] == 0 and r['n_bits'] == len(g['bits'])
        assert np.max(np.abs(r['papr_no_cp']['papr_per_symbol'] - g[f'{tag}_no_cp_db'])) < DB_TOL


def test_papr_symbols_edge_cases():
    import torch
    from lte_b200 import LinkEngine
    eng = LinkEngine(128, 76, 9, 2, 1.92e6)
    x = torch.zeros((2, 3 * eng.L + 5), dtype=torch.complex64, device='cuda')     # ragged tail is ignored
    x[1, eng.L + eng.cp + 7] = 3 + 4j                                             # one spike in symbol 1
    x[1, 2 * eng.L + 1] = 1.0                                                     # inside the CP of symbol 2
    db, pm = eng.papr_symbols(x, want_peak_mean=True)
    assert db.shape == (2, 3)
    want = np.zeros((2, 3))
    want[1, 1] = 10 * np.log10(128.0)
    assert np.allclose(db.cpu().numpy(), want, atol=1e-4)
    assert float(pm[1, 1, 0]) == 25.0 and float(pm[1, 2, 0]) == 0.0
    db_cp, _ = eng.papr_symbols(x, include_cp=True)
    assert abs(float(db_cp[1, 2]) - 10 * np.log10(eng.L)) < 1e-4
    e, _ = eng.papr_symbols(torch.zeros((1, 5), dtype=torch.complex64, device='cuda'))
    assert e.shape == (1, 0)
    h = eng.histogram(torch.tensor([-5.0, 0.2, 0.2, 99.0], device='cuda'), 0.0, 0.1, 10)
    assert h.cpu().tolist() == [1, 0, 2, 0, 0, 0, 0, 0, 0, 1]


def test_sweep_is_batch_invariant_and_sc_fdm_lowers_papr():
    from lte_b200 import LinkEngine
    from lte_b200.papr import papr_sweep
    num = O.Numerology(10.0, 15.0, '16-QAM')
    eng = LinkEngine(num.N, num.Nc, num.cp_length, num.bits_per_symbol, num.fs)
    a = papr_sweep(eng, 300, symbols_per_stream=14, seed=5, batch_streams=300, return_values=True)
    b = papr_sweep(eng, 300, symbols_per_stream=14, seed=5, batch_streams=64)
    assert a['count'] == 300 * 14 and np.array_equal(a['hist'], b['hist'])
    # two "ranks" sharing the range see the same union of symbols
    h0 = papr_sweep(eng, 300, symbols_per_stream=14, seed=5, rank=0, world=2)['hist']
    h1 = papr_sweep(eng, 300, symbols_per_stream=14, seed=5, rank=1, world=2)['hist']
    assert np.array_equal(h0 + h1, a['hist'])
    vals = a['values'].cpu().numpy().astype(np.float64)
    assert np.array_equal(a['hist'], O.papr_histogram(vals, a['lo'], a['step'], len(a['hist'])))
    thr = a['thresholds_db']
    assert np.max(np.abs(a['ccdf'] - O.ccdf(vals, thr))) < 2e-3       # bin-edge ties only
    s = papr_sweep(eng, 300, symbols_per_stream=14, seed=5, sc_fdm=True)
    assert s['mean_db'] < a['mean_db'] - 0.3
    # textbook anchor: OFDM PAPR exceeds 8 dB for a sizeable fraction of symbols, rarely 12 dB
    assert 0.05 < O.ccdf(vals, [8.0])[0] < 0.9 and O.ccdf(vals, [12.5])[0] < 0.01


def test_collect_papr_fused_mode():
    from config import LTEConfig
    from core.ofdm_system import OFDMSystem
    sysm = OFDMSystem(LTEConfig(5.0, 15.0, '16-QAM'), 'awgn', rng='philox', seed=3)
    nbits = 4 * 249 * 2
    r = sysm.collect_papr_for_all_modulations(nbits, 25)
    assert set(r) == {'QPSK_OFDM', 'QPSK_SC-FDM', '16-QAM_OFDM', '16-QAM_SC-FDM'}
    assert len(r['QPSK_OFDM']) == 25 * 4 and len(r['16-QAM_OFDM']) == 25 * 2
    assert r['QPSK_SC-FDM'].mean() < r['QPSK_OFDM'].mean()
    assert sysm.config.modulation == '16-QAM' and sysm.config.bits_per_symbol == 4     # restored
    # faithful mode consumes np.random like the reference and yields the same number of samples
    np.random.seed(11)
    sysn = OFDMSystem(LTEConfig(5.0, 15.0, '16-QAM'), 'awgn')
    rn = sysn.collect_papr_for_all_modulations(nbits, 2)
    assert len(rn['QPSK_OFDM']) == 2 * 4 and len(rn['16-QAM_SC-FDM']) == 2 * 2
