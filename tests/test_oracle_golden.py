"""Pins oracle/lte_oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from cases import BIG_RX_STRIDE, SIMO_CASES, SISO_CASES
from helpers import (golden_bits, golden_bits_rx, load_golden, oracle_simo, oracle_siso,
                     reference_draws, rel_err)
from oracle import lte_oracle as O

TOL64 = 1e-12     # complex128 fixtures
TOL32 = 2e-7      # fixtures stored as complex64


def test_tables_match_reference():
    g = load_golden('tables')
    for bw in (1.25, 2.5, 5.0, 10.0, 15.0, 20.0):
        for cp in ('normal', 'extended'):
            n = O.Numerology(bw, 15.0, 'QPSK', cp)
            assert [n.N, n.Nc, n.cp_length, n.fs] == list(g[f'num_{bw}_{cp}'])
        d, p = O.grid_indices(n.N, n.Nc)
        assert np.array_equal(d, g[f'data_idx_{bw}'])
        assert np.array_equal(p, g[f'pilot_idx_{bw}'])
    n = O.Numerology(3.0, 15.0, 'QPSK')
    assert [n.N, n.Nc, n.cp_length, n.fs] == list(g['num_3.0_normal'])
    n = O.Numerology(5.0, 7.5, 'QPSK', 'extended')
    assert [n.N, n.Nc, n.cp_length, n.fs] == list(g['num_5.0_7.5_extended'])
    for cell in range(4):
        assert np.array_equal(O.pilots(cell, 200), g[f'pilots_cell{cell}'])
        assert np.array_equal(O.pilots(cell, 50), g[f'pilots_cell{cell}'][:50])   # prefix property


@pytest.mark.parametrize('mod', ['QPSK', '16-QAM', '64-QAM'])
def test_qam_map_demap_match_reference(mod):
    g = load_golden('tables')
    assert np.array_equal(O.constellation(mod), g[f'const_{mod}'])
    assert np.array_equal(O.qam_map(g[f'mapbits_{mod}'], mod), g[f'mapsyms_{mod}'])
    assert np.array_equal(O.qam_demap(g[f'demapsyms_{mod}'], mod), g[f'demapbits_{mod}'])


@pytest.mark.parametrize('case', SISO_CASES, ids=lambda c: c['name'])
def test_siso_matches_reference(case):
    g = load_golden(case['name'])
    bits = golden_bits(g)
    tol = TOL32 if case.get('big') else TOL64
    for snr in case['snrs']:
        o = oracle_siso(case, bits, snr)
        assert o['errors'] == int(g[f'errors_{snr}'])
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        assert abs(o['papr_db'] - float(g[f'papr_db_{snr}'])) < 1e-9
        if snr == case['full_snr']:
            assert rel_err(o['signal_tx'], g['signal_tx']) < tol
            assert rel_err(o['signal_rx'], g['signal_rx']) < tol
            assert rel_err(o['symbols_rx'], g['symbols_rx']) < tol


@pytest.mark.parametrize('case', SIMO_CASES, ids=lambda c: c['name'])
def test_simo_matches_reference(case):
    g = load_golden(case['name'])
    bits = golden_bits(g)
    big = case.get('big')
    tol = TOL32 if big else TOL64
    for snr in case['snrs']:
        o = oracle_simo(case, bits, snr)
        assert o['errors'] == int(g[f'errors_{snr}'])
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        if snr == case['full_snr']:
            assert rel_err(o['signal_tx'], g['signal_tx']) < tol
            rx = o['signal_rx'][:, ::BIG_RX_STRIDE] if big else o['signal_rx']
            assert rel_err(rx, g['signal_rx']) < tol
            assert rel_err(o['symbols_combined'], g['symbols_combined']) < tol
            assert rel_err(o['H'][:, ::14, :], g['H']) < tol


def test_injected_draws_equal_reference_draws():
    """The replay interface (explicit phases / unit normals) reproduces the
    reference's own global-RNG draws."""
    case = SIMO_CASES[0]
    g = load_golden(case['name'])
    bits = golden_bits(g)
    snr = case['full_snr']
    n = len(g['signal_tx'])
    phases, z = reference_draws(case, n, case['R'])
    o = oracle_simo(case, bits, snr, phases=phases, z=z)
    assert o['errors'] == int(g[f'errors_{snr}'])
    assert rel_err(o['signal_rx'], g['signal_rx']) < TOL64


from cases import SFBC_CASES  # noqa: E402


@pytest.mark.parametrize('case', SFBC_CASES, ids=lambda c: c['name'])
def test_sfbc_matches_reference(case):
    from helpers import numerology
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    for snr in case['snrs']:
        o = O.simulate_sfbc(bits, snr, num, case['R'], case['ch'], case['prof'], 2.0, case['v'])
        assert o['errors'] == int(g[f'errors_{snr}'])
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        assert rel_err(o['channel_matrix'], g[f'channel_matrix_{snr}']) < TOL64
        assert abs(o['papr_db_tx0'] - g[f'papr_{snr}'][0]) < 1e-9
        assert abs(o['papr_db_tx1'] - g[f'papr_{snr}'][1]) < 1e-9


from cases import SM_CASES  # noqa: E402


@pytest.mark.parametrize('case', SM_CASES, ids=lambda c: c['name'])
def test_spatial_multiplexing_matches_reference(case):
    from helpers import numerology
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    for snr in case['snrs']:
        o = O.simulate_sm(bits, num, case['T'], case['R'], case['rank'], case['det'], snr, case['ch'], case['prof'],
                          case['v'], 2.0, global_seed=case['gseed'])
        assert o['errors'] == int(g[f'errors_{snr}'])
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        assert [o['rank'], o['pmi']] == list(g[f'rank_pmi_{snr}'])
        assert rel_err(o['W'], g[f'W_{snr}']) < TOL64
        assert rel_err(o['channel_matrix'], g[f'channel_matrix_{snr}']) < TOL64


from cases import BF_CASES  # noqa: E402


@pytest.mark.parametrize('case', BF_CASES, ids=lambda c: c['name'])
def test_beamforming_matches_reference(case):
    """OFDMSimulator.simulate_beamforming (core/ofdm_core.py:2260-2477)."""
    from helpers import numerology
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    for snr in case['snrs']:
        o = O.simulate_beamforming(bits, snr, num, case['T'], case['R'], case['upd'], global_seed=case['gseed'])
        assert o['errors'] == int(g[f'errors_{snr}'])
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        assert rel_err(o['channel_matrix'], g[f'channel_matrix_{snr}']) < TOL64
        assert o['pmi_history'] == list(g[f'pmi_history_{snr}'])
        assert abs(o['beamforming_gain_db'] - g[f'gain_unique_{snr}'][0]) < 1e-9
        assert o['unique_pmis'] == int(g[f'gain_unique_{snr}'][1])


def test_beamforming_helpers_match_reference_tables():
    """CQI table (core/csi_feedback.py:106-137) and update period (core/beamforming_precoder.py:231-263)."""
    edges = [-7.0, -6.0, -4.1, -4.0, 0.0, 1.99, 2.0, 11.0, 21.99, 22.0, 40.0]
    assert [O.sinr_to_cqi(x) for x in edges] == [0, 1, 1, 2, 4, 4, 5, 9, 14, 15, 15]
    assert O.beamforming_update_period(0.0) == 100
    assert O.beamforming_update_period(3.0) == 48           # values printed by the reference's AdaptiveBeamforming
    assert O.beamforming_update_period(120.0) == 1
