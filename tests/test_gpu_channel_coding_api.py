"""The reference's module-level channel-coding API (core/channel_coding/*.py) on the GPU kernels, checked
against the vectors the reference's own functions produced (tests/golden/coding_tables.npz)."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu
T = load_golden('coding_tables')
QPP = {int(k): (int(f1), int(f2)) for k, f1, f2 in T['qpp']}


@pytest.mark.parametrize('n', [1, 24, 100, 1001])
def test_crc(n):
    from core.channel_coding import crc
    b = T[f'crc_in_{n}']
    assert np.array_equal(crc.calculate_crc24a(b), T[f'crc24a_{n}'])
    assert np.array_equal(crc.calculate_crc24b(b), T[f'crc24b_{n}'])
    full = crc.attach_crc24a(b)
    assert len(full) == n + 24 and crc.check_crc24a(full) and not crc.check_crc24b(full) or n == 0
    bad = full.copy()
    bad[0] ^= 1
    assert not crc.check_crc24a(bad) and not crc.check_crc24a(b[:5])
    c16 = crc.attach_crc16(b)
    assert len(c16) == n + 16 and crc.check_crc16(c16)
    assert np.array_equal(crc.calculate_crc16(b), O.crc24(b, 0x1021)[-16:]) or True      # 16-bit register checked below
    reg = 0
    for bit in b:
        top = ((reg >> 15) & 1) ^ int(bit)
        reg = (reg << 1) & 0xFFFF
        if top:
            reg ^= 0x1021
    assert np.array_equal(crc.calculate_crc16(b), [(reg >> (15 - i)) & 1 for i in range(16)])


@pytest.mark.parametrize('B', [40, 41, 100, 6144, 6145, 13000, 20011])
def test_segmentation(B):
    from core.channel_coding import desegment_code_blocks, get_segmentation_info, segment_code_blocks
    tb = T[f'seg_in_{B}']
    blocks, meta = segment_code_blocks(tb)
    assert [len(b) for b in blocks] == list(T[f'seg_sizes_{B}']) == meta['block_sizes']
    assert np.array_equal(np.concatenate(blocks), T[f'seg_out_{B}'])
    assert np.array_equal(desegment_code_blocks(blocks, meta), tb)
    assert meta['original_size'] == B and meta['segmented'] == (B > 6144) and meta['num_blocks'] == len(blocks)
    info = get_segmentation_info(B)
    assert info['block_sizes'] == meta['block_sizes'] and info['total_coded_bits'] == sum(3 * k + 12 for k in meta['block_sizes'])


@pytest.mark.parametrize('K', [40, 104, 512, 6144])
def test_encoder_rate_matching_and_interleavers(K):
    from core.channel_coding import (qpp_deinterleave, qpp_interleave, rate_dematching_turbo, rate_match_turbo,
                                     sub_block_deinterleaver, sub_block_interleaver, turbo_encode)
    from core.channel_coding.turbo_encoder import rsc_encode
    u = T[f'enc_in_{K}']
    enc = turbo_encode(u)
    assert np.array_equal(enc, T[f'enc_out_{K}'])
    assert np.array_equal(rate_match_turbo(enc, len(enc), K, rv_idx=0), T[f'rm_out_{K}'])
    assert np.allclose(rate_dematching_turbo(T[f'dm_in_{K}'], K, rv_idx=0), T[f'dm_out_{K}'], rtol=1e-6, atol=1e-6)
    pi = O.qpp_indices(K, *QPP[K])
    x = np.arange(K) % 251
    assert np.array_equal(qpp_interleave(x, K), x[pi]) and np.array_equal(qpp_deinterleave(qpp_interleave(x, K), K), x)
    d = (np.arange(K + 6) * 7) % 2
    v = sub_block_interleaver(d)
    assert np.array_equal(v, d[O.sub_block_permutation(K + 6)]) and np.array_equal(sub_block_deinterleaver(v, K + 6), d)
    s, p = rsc_encode(u[:37])
    so, po = O.rsc_encode(u[:37])
    assert np.array_equal(s, so) and np.array_equal(p, po)
    # other redundancy versions and lengths: wrap-around of the circular buffer, repetition accumulates on the way back
    for rv, E in ((1, 2 * K), (2, 3 * K + 12), (3, 4 * K)):
        rm = rate_match_turbo(enc, E, K, rv_idx=rv)
        cbuf = np.concatenate([np.where(t >= 0, enc[np.maximum(t, 0)], 0) for t in [np.asarray(
            __import__('core.channel_coding.rate_matching', fromlist=['x'])._circular_buffer_source(K))]])
        N = len(cbuf)
        start = [0, N // 4, N // 2, 3 * N // 4][rv]
        assert np.array_equal(rm, cbuf[(start + np.arange(E)) % N])
    with pytest.raises(ValueError):
        turbo_encode(u[:39])
    with pytest.raises(ValueError):
        rate_match_turbo(enc[:-1], 10, K)


@pytest.mark.parametrize('K', [40, 104, 512])
def test_turbo_decode_function(K):
    from core.channel_coding import LogMAPDecoder, turbo_decode
    llr = T[f'dec_llr_{K}']
    l32 = llr.astype(np.float32).astype(float)
    one = turbo_decode(llr, K, num_iterations=1)
    assert np.sum(one != O.turbo_decode(l32, K, *QPP[K], num_iterations=1)) <= 2
    if K == 40:                                    # this vector converges: the reference's own decisions, bit for bit
        assert np.array_equal(turbo_decode(llr, K, num_iterations=8), T[f'dec_out_{K}'])
    Ls = np.concatenate([l32[0:3 * K:3], l32[3 * K:3 * K + 3]])
    Lp = np.concatenate([l32[1:3 * K:3], l32[3 * K + 3:3 * K + 6]])
    La = np.concatenate([0.3 * np.sin(np.arange(K)), np.zeros(3)]).astype(np.float32).astype(float)
    bits, ext = LogMAPDecoder().decode(Ls, Lp, La, return_extrinsic=True)
    wb, we = O.maxlog_bcjr(Ls, Lp, La, extrinsic=True)
    assert np.allclose(ext, we, rtol=1e-4, atol=2e-4)
    assert np.sum(bits != wb) <= np.sum(np.abs(we + La + Ls) < 1e-3)
    _, ap = LogMAPDecoder().decode(Ls, Lp, None, return_extrinsic=False)
    assert np.allclose(ap, O.maxlog_bcjr(Ls, Lp, np.zeros(K + 3), extrinsic=False)[1], rtol=1e-4, atol=2e-4)
    with pytest.raises(ValueError):
        LogMAPDecoder().decode(Ls, Lp, np.ones(K + 3))


def test_exact_log_map_mode():
    """set_decoder_mode(False): max*(a, b) = log(e^a + e^b) in the kernel, against the reference's vectors."""
    from core.channel_coding import LogMAPDecoder, turbo_decode
    from core.channel_coding import turbo_decoder as td
    Ls, Lp, La = T['bcjr_logmap_in']
    try:
        td.set_decoder_mode(False)
        assert td.max_star(0.0, 0.0) == pytest.approx(np.log(2.0)) and td.max_star(-np.inf, 1.5) == 1.5
        _, ext = LogMAPDecoder().decode(Ls, Lp, La, return_extrinsic=True)
        assert np.allclose(ext, T['bcjr_logmap_ext'], rtol=1e-4, atol=1e-4)
        got = turbo_decode(T['dec_llr_40'], 40, num_iterations=8)
        assert np.array_equal(got, T['dec_out_logmap_40'])
        one = turbo_decode(T['dec_llr_104'], 104, num_iterations=1)
        l32 = T['dec_llr_104'].astype(np.float32).astype(float)
        assert np.sum(one != O.turbo_decode(l32, 104, *QPP[104], num_iterations=1, logmap=True)) <= 2
    finally:
        td.set_decoder_mode(True)
    _, ext = LogMAPDecoder().decode(Ls, Lp, La, return_extrinsic=True)
    assert np.allclose(ext, T['bcjr_maxlog_ext'], rtol=1e-4, atol=1e-4) and td.max_star(0.0, 0.0) == 0.0
