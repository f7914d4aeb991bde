"""Pins the oracle's per-symbol PAPR / histogram / CCDF restatement against the
reference's OFDMSystem (tests/golden/make_golden.py: papr_case).  CPU only."""
import numpy as np
import pytest

from cases import PAPR_CASES
from helpers import load_golden
from oracle import lte_oracle as O


@pytest.mark.parametrize('case', PAPR_CASES, ids=[c['name'] for c in PAPR_CASES])
def test_papr_per_symbol_matches_reference(case):
    g = load_golden(case['name'])
    num = O.Numerology(case['bw'], 15.0, case['mod'])
    for tag, flag in (('ofdm', False), ('scfdm', True)):
        sig, _ = O.modulate_stream(g['bits'], num, sc_fdm=flag)
        db, _, _ = O.papr_per_symbol_db(sig, num)
        assert np.max(np.abs(db - g[f'{tag}_no_cp_db'])) < 1e-9
        db_cp, peak, mean = O.papr_per_symbol_db(sig, num, include_cp=True)
        assert np.max(np.abs(db_cp - g[f'{tag}_cp_db'])) < 1e-9
        assert np.allclose(peak, g[f'{tag}_cp_peak'], rtol=1e-11, atol=0)
        assert np.allclose(mean, g[f'{tag}_cp_avg'], rtol=1e-11, atol=0)
        assert np.allclose(10 ** (db / 10), O.papr_per_symbol_no_cp(sig, num), rtol=1e-12)
        s_db, s_lin = O.papr(sig)
        assert abs(s_db - g[f'{tag}_stream'][0]) < 1e-9 and abs(s_lin / g[f'{tag}_stream'][1] - 1) < 1e-11


def test_sc_fdm_lowers_papr_of_cfg2():
    """The point of config 2's comparison: the DFT-precoded waveform peaks lower on average."""
    g = load_golden('papr_cfg2_10mhz_16qam')
    assert g['scfdm_no_cp_db'].mean() < g['ofdm_no_cp_db'].mean()


def test_all_zero_symbol_reports_zero_db():
    num = O.Numerology(1.25, 15.0, 'QPSK')
    db, peak, mean = O.papr_per_symbol_db(np.zeros(2 * num.L, dtype=complex), num)
    assert np.array_equal(db, [0.0, 0.0]) and not peak.any() and not mean.any()


def test_histogram_and_ccdf_rules():
    v = np.array([-1.0, 0.0, 0.049, 0.051, 3.31, 19.99, 20.0, 55.0])
    h = O.papr_histogram(v, 0.0, 0.05, 400)
    assert h.sum() == len(v) and h[0] == 3 and h[1] == 1 and h[66] == 1 and h[399] == 3
    thr = np.array([-2.0, 0.0, 3.31, 60.0])
    assert np.allclose(O.ccdf(v, thr), [1.0, 6 / 8, 3 / 8, 0.0])
    from lte_b200_papr_host import ccdf_from_hist
    edges, c = ccdf_from_hist(h, 0.0, 0.05)
    assert c[-1] == 0 and abs(c[0] - 5 / 8) < 1e-12 and np.all(np.diff(c) <= 0)
    assert abs(edges[0] - 0.05) < 1e-12 and len(edges) == 400
