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
