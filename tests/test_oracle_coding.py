"""Oracle of the coded chain against vectors produced by the reference's
core/channel_coding/*.py and OFDMSimulator.simulate_siso_coded (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from cases import CODED_CASES
from helpers import golden_bits, golden_bits_rx, load_golden, numerology, rel_err
from oracle import lte_oracle as O

T = load_golden('coding_tables')
QPP = {int(k): (int(f1), int(f2)) for k, f1, f2 in T['qpp']}


def test_qpp_table_and_block_sizes():
    assert sorted(QPP) == O.TURBO_K and len(QPP) == 188
    for K in (40, 512, 6144):                       # a permutation for every size we touch
        assert np.array_equal(np.sort(O.qpp_indices(K, *QPP[K])), np.arange(K))


@pytest.mark.parametrize('n', [1, 24, 100, 1001])
def test_crc24(n):
    b = T[f'crc_in_{n}']
    assert np.array_equal(O.crc24(b, O.CRC24A_POLY), T[f'crc24a_{n}'])
    assert np.array_equal(O.crc24(b, O.CRC24B_POLY), T[f'crc24b_{n}'])


@pytest.mark.parametrize('B', [40, 41, 100, 6144, 6145, 13000, 20011])
def test_segmentation(B):
    tb = T[f'seg_in_{B}']
    blocks = O.segment_code_blocks(tb)
    assert [len(b) for b in blocks] == list(T[f'seg_sizes_{B}'])
    assert np.array_equal(np.concatenate(blocks), T[f'seg_out_{B}'])
    assert np.array_equal(O.desegment_code_blocks(blocks, B), tb)


@pytest.mark.parametrize('K', [40, 104, 512, 6144])
def test_turbo_encoder_and_rate_matching(K):
    enc = O.turbo_encode(T[f'enc_in_{K}'], *QPP[K])
    assert np.array_equal(enc, T[f'enc_out_{K}'])
    t = O.rate_match_table(K)
    assert np.array_equal(np.where(t >= 0, enc[np.maximum(t, 0)], 0), T[f'rm_out_{K}'])
    d = O.rate_dematch_table(K)
    l = T[f'dm_in_{K}']
    assert np.array_equal(np.where(d >= 0, l[np.maximum(d, 0)], 0.0), T[f'dm_out_{K}'])


@pytest.mark.parametrize('K', [40, 104, 512])
def test_turbo_decoder(K):
    got = O.turbo_decode(T[f'dec_llr_{K}'], K, *QPP[K], num_iterations=8)
    assert np.array_equal(got, T[f'dec_out_{K}'])


@pytest.mark.parametrize('mod', ['QPSK', '16-QAM', '64-QAM'])
def test_soft_demapper(mod):
    got = O.soft_demap(T['llr_sym'], T['llr_nv'], mod)
    assert np.allclose(got, T[f'llr_{mod}'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('case', CODED_CASES, ids=lambda c: c['name'])
def test_simulate_siso_coded_matches_reference(case):
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    for snr in case['snrs']:
        o = O.simulate_siso_coded(bits, snr, num, QPP, case['ch'], case['prof'], 2.0, case['v'])
        assert rel_err(o['symbols_rx'], g[f'symbols_rx_{snr}']) < 1e-12
        assert rel_err(o['H_estimate'], g[f'H_{snr}']) < 1e-12
        assert np.array_equal(o['bits_rx'], golden_bits_rx(g, snr))
        assert o['errors'] == int(g[f'errors_{snr}'])
        crc, clen, papr_db, nvm = g[f'scalars_{snr}']
        assert o['crc_pass'] == bool(crc) and o['coded_bits_length'] == int(clen)
        assert abs(o['papr_db'] - papr_db) < 1e-9 and abs(o['noise_var_mean'] - nvm) < 1e-12 * max(1, nvm)
    assert rel_err(o['signal_tx'], g['signal_tx']) < 1e-12


def test_bcjr_modes_match_the_reference():
    """One LogMAPDecoder.decode pass in both modes, and the exact log-MAP decoder's decisions."""
    Ls, Lp, La = T['bcjr_logmap_in']
    assert np.allclose(O.maxlog_bcjr(Ls, Lp, La)[1], T['bcjr_maxlog_ext'], rtol=1e-12, atol=1e-12)
    assert np.allclose(O.maxlog_bcjr(Ls, Lp, La, logmap=True)[1], T['bcjr_logmap_ext'], rtol=1e-10, atol=1e-10)
    for K in (40, 104):
        assert np.array_equal(O.turbo_decode(T[f'dec_llr_{K}'], K, *QPP[K], logmap=True), T[f'dec_out_logmap_{K}'])
