"""NumPy fp64 restatement of the *spectral* fading path (csrc/spectral.cu) -- test infrastructure.

For low Doppler the Jakes process of every (antenna, tap) is linear over one OFDM symbol to better than
5e-7 (the economised K = 1 fit of csrc/tdl.cuh), h(m) = c0 + c1 tau(m), tau(m) = m - (L - 1)/2.  The
time-domain tapped delay line of the reference (core/rayleighchannel.py:44-58) followed by CP strip and
fft / sqrt(N) (core/lte_receiver.py:444-491) then collapses, bin by bin, to

    Y_r[k] = sum_t e_t[k] { (c0 + d_t c1) X[k] + c1 (G[k] - N T_t[k]) },   e_t[k] = exp(-2 pi j k d_t / N)

with X the transmitted grid, G = fft((n - n_c) u[n]) / sqrt(N) the spectrum of the ramp-weighted useful
samples u (ONE transform per OFDM symbol, independent of antenna and tap), and T_t[k] the partial DFT of
the last d_t samples of u: the samples whose delayed copy arrives through the cyclic prefix and therefore
sees the ramp N samples earlier.  Only the CP region of the received stream -- which enters nothing but
the noise power (core/channel.py:216-218) -- is still evaluated sample by sample.
"""
import numpy as np

from oracle import lte_oracle as O


def linear_coefficients(num, S, fD, delays, gains, phases):
    """c0, c1 [R, taps, S] of h(m) ~ c0 + c1 (m_local - (L-1)/2) per OFDM symbol: the economised
    (Chebyshev) linear fit of csrc/tdl.cuh:61-67 around the symbol centre."""
    R, T = phases.shape[0], len(delays)
    L = num.L
    alpha = 2 * np.pi * np.arange(1, 17) / 16
    w = 2 * np.pi * fD * np.cos(alpha) / num.fs               # rad / sample
    mc = np.arange(S) * L + 0.5 * (L - 1)
    X = w * 0.5 * L
    c0 = np.zeros((R, T, S), complex)
    c1 = np.zeros((R, T, S), complex)
    for r in range(R):
        for t in range(T):
            ph = np.exp(1j * (w[None, :] * mc[:, None] + phases[r, t][None, :]))      # [S, 16]
            g = gains[t] * np.sqrt(2 / 16)
            c0[r, t] = g * (ph * (1 - 0.25 * X ** 2)[None, :]).sum(1)
            c1[r, t] = g * (ph * (1j * w)[None, :]).sum(1)
    return c0, c1


def spectral_rx(signal_tx, num, fD, delays, gains, phases, kept=None):
    """-> (Y [R, S, N] on the bins `kept` (zero elsewhere), power [R] = sum |faded|^2 over the stream
    with the useful part taken over the kept bins only)."""
    N, cp, L = num.N, num.cp_length, num.L
    S = len(signal_tx) // L
    x = np.asarray(signal_tx).reshape(S, L)
    u = x[:, cp:]
    R = phases.shape[0]
    c0, c1 = linear_coefficients(num, S, fD, delays, gains, phases)
    k = np.arange(N)
    if kept is None:
        kept = k
    nc = 0.5 * (L - 1) - cp
    Xg = np.fft.fft(u, axis=-1) / np.sqrt(N)
    G = np.fft.fft((np.arange(N) - nc)[None, :] * u, axis=-1) / np.sqrt(N)
    Y = np.zeros((R, S, N), complex)
    power = np.zeros(R)
    for t, d in enumerate(delays):
        e = np.exp(-2j * np.pi * k * d / N)
        Tt = np.zeros((S, N), complex)
        for i in range(1, d + 1):                                    # tail samples u[N - i]
            Tt += u[:, N - i][:, None] * np.exp(2j * np.pi * k * i / N)[None, :]
        Tt /= np.sqrt(N)
        for r in range(R):
            a = (c0[r, t] + d * c1[r, t])[:, None]
            Y[r] += e[None, :] * (a * Xg + c1[r, t][:, None] * (G - N * Tt))
    mask = np.zeros(N, bool)
    mask[kept] = True
    Y[:, :, ~mask] = 0
    power += (np.abs(Y) ** 2).sum(axis=(1, 2))
    # cyclic-prefix region, sample by sample (previous symbol's tail for m_local < d; zeros before the stream)
    flat = np.concatenate([np.zeros(max(delays) if len(delays) else 0, complex), np.asarray(signal_tx)])
    off = max(delays) if len(delays) else 0
    tau = np.arange(cp) - 0.5 * (L - 1)
    for s in range(S):
        for r in range(R):
            y = np.zeros(cp, complex)
            for t, d in enumerate(delays):
                xs = flat[off + s * L - d: off + s * L - d + cp]
                y += (c0[r, t, s] + c1[r, t, s] * tau) * xs
            power[r] += (np.abs(y) ** 2).sum()
    return Y, power


def time_domain_rx(signal_tx, num, fD, delays, gains, phases):
    """The oracle's own path: exact Jakes TDL per antenna, CP strip, fft / sqrt(N), stream power."""
    R = phases.shape[0]
    Y, power = [], []
    for r in range(R):
        faded = O.rayleigh_filter(signal_tx, num.fs, fD, delays, gains, phases[r])
        Y.append(O.rx_fft_stream(faded, num))
        power.append((np.abs(faded) ** 2).sum())
    return np.stack(Y), np.array(power)
