"""Known-answer tests the reference itself asserts on this path, restated against the oracle:
Alamouti encode identities and perfect-channel decode (reference test/test_alamouti_unit.py:49-52,
:119), noisy SER bound (:137-180), fft(ifft(x)) round trip (core/parallel_processing.py:263-270)."""
import numpy as np

from oracle import lte_oracle as O


def test_alamouti_encode_identities():
    s = np.array([1 + 1j, -1 + 1j, 3 - 1j, -2 - 2j])
    tx0, tx1 = O.sfbc_encode(s)
    for i in (0, 2):
        assert np.isclose(tx0[i], s[i]) and np.isclose(tx1[i], s[i + 1])
        assert np.isclose(tx0[i + 1], -np.conj(s[i + 1])) and np.isclose(tx1[i + 1], np.conj(s[i]))


def test_alamouti_perfect_channel_decode():
    s = np.array([1 + 1j, -1 + 1j])
    h0, h1 = 1.0 + 0j, 0.0 + 1j
    tx0, tx1 = O.sfbc_encode(s)
    rx = h0 * tx0 + h1 * tx1
    dec = O.sfbc_decode(rx, np.full(2, h0), np.full(2, h1))
    assert np.max(np.abs(dec - s)) < 1e-10


def test_alamouti_noisy_ser_below_ten_percent():
    np.random.seed(42)
    c = O.constellation('QPSK')
    s = c[np.random.randint(0, 4, 1000)]
    h0 = (np.random.randn() + 1j * np.random.randn()) / np.sqrt(2)
    h1 = (np.random.randn() + 1j * np.random.randn()) / np.sqrt(2)
    tx0, tx1 = O.sfbc_encode(s)
    sigma = np.sqrt(10 ** (-10 / 10) / 2)
    rx = h0 * tx0 + h1 * tx1 + sigma * (np.random.randn(1000) + 1j * np.random.randn(1000))
    dec = O.sfbc_decode(rx, np.full(1000, h0), np.full(1000, h1))
    det = c[O.qam_demap_indices(dec, 'QPSK')]
    assert np.mean(det != s) < 0.10


def test_fft_ifft_round_trip():
    x = np.random.RandomState(0).randn(2048) + 1j * np.random.RandomState(1).randn(2048)
    num = O.Numerology(20.0, 15.0, 'QPSK')
    td = O.ofdm_modulate_grid(x[None, :], num)
    back = O.rx_fft_stream(td.reshape(-1), num)
    assert np.allclose(back[0], x, atol=1e-10)


def test_mimo_detector_toy_case():
    """reference core/mimo_detector.py:387-404: 2x2 MMSE on a fixed H, ||s - s_hat|| < 1."""
    H = np.array([[1.0 + 0.5j, 0.3 - 0.2j], [0.2 + 0.1j, 0.9 - 0.3j]])
    s = np.array([1 + 1j, -1 + 1j]) / np.sqrt(2)
    y = H @ s + 0.1 * (np.array([0.3, -0.2]) + 1j * np.array([0.1, 0.25]))
    for det in ('MMSE', 'ZF'):
        assert np.linalg.norm(s - O.mimo_detect(y, H, 0.01, det)) < 1.0
    c = O.constellation('QPSK')
    assert np.allclose(O.mimo_detect(y, H, 0.01, 'SIC', c), s)
    assert O.mimo_detect(y, H[:, :1], 0.01, 'MRC').shape == (1,)


def test_layer_mapper_round_trips():
    """reference core/layer_mapper.py:172-219."""
    for rank in (1, 2, 3, 4):
        for n in (12, 13, 999):
            x = np.arange(n) + 1j * np.arange(n)
            lay = O.layer_map(x, rank)
            assert lay.shape == (rank, -(-n // rank))
            assert np.array_equal(O.layer_demap(lay, n), x)
    lay = O.layer_map(np.arange(6), 2)
    assert np.array_equal(lay, [[0, 2, 4], [1, 3, 5]])


def test_rank_adaptation_rules():
    """reference core/rank_adaptation.py:275-333: RI range, W shape, SNR < 5 dB => RI = 1."""
    H = (np.random.RandomState(0).randn(4, 4) + 1j * np.random.RandomState(1).randn(4, 4)) / np.sqrt(8)
    ri, pmi, W = O.rank_feedback(H, 4, 4, 20.0)
    assert 1 <= ri <= 4 and W.shape == (4, ri) and 0 <= pmi < len(O.codebook(4, ri))
    assert O.rank_feedback(H, 4, 4, 3.0)[0] == 1
    assert O.rank_feedback(H, 4, 4, 8.0)[0] <= 2
    for T in (2, 4, 8):
        for r in range(1, min(T, 4) + 1):
            for W in O.codebook(T, r):
                assert W.shape == (T, r)
