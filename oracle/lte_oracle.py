"""CPU oracle for the LTE simulate-and-count-BER link chain.

TEST INFRASTRUCTURE ONLY.  This module is a vectorised NumPy (fp64) restatement
of the algorithm in the reference (Darioxavierl/OFDM-LTE, pure Python/NumPy).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product path
(``ofdm-lte_b200/``) never does and fails loudly without its CUDA library.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference`` in the build container, runs it, and commits
its outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those vectors (signals <= 1e-12 relative, bits and
error counts identical).  The reference's own asserting tests on this path
(Alamouti identities ``test/test_alamouti_unit.py:49-52,119``; the 2x2 MMSE toy
``core/mimo_detector.py:387-404``; layer-mapper round trips
``core/layer_mapper.py:172-219``) are restated in ``tests/test_oracle_kat.py``.

Third-party arithmetic the reference relies on and that is restated by calling
the same library: ``numpy.fft`` (pocketfft, unitary scaling applied by the
caller), ``numpy.linalg.inv/pinv`` (LAPACK) and the legacy ``numpy.random``
MT19937 stream (versions unpinned in ``requirements.txt``; this image has
numpy 2.3.5).

Every function cites the reference file:line it follows (paths relative to the
reference root).  All faithfulness quirks are kept: natural-binary raster QAM,
un-shifted FFT-bin grid, double dB->linear conversion of tap gains, power-2
Jakes, per-14-symbol channel hold, 1e-6 / 1e-10 regularisers.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# L0: numerology and tables                                   (config.py:11-60)
# ----------------------------------------------------------------------------
LTE_PROFILES = {1.25: (76, 128), 2.5: (150, 256), 5.0: (300, 512),
                10.0: (600, 1024), 15.0: (900, 2048), 20.0: (1200, 2048)}
CP_US = {'normal': 4.7, 'extended_15khz': 16.6, 'extended_7.5khz': 33.0}
BITS_PER_SYMBOL = {'QPSK': 2, '16-QAM': 4, '64-QAM': 6}
ITU = {
    'Pedestrian_A': ([0.0, 0.11, 0.19, 0.41], [0.0, -9.7, -19.2, -22.8]),
    'Pedestrian_B': ([0.0, 0.2, 0.8, 1.2, 2.3, 3.7], [0.0, -0.9, -4.9, -8.0, -7.8, -23.9]),
    'Vehicular_A': ([0.0, 0.31, 0.71, 1.09, 1.73, 2.51], [0.0, -1.0, -9.0, -10.0, -15.0, -20.0]),
    'Vehicular_B': ([0.0, 0.3, 0.7, 1.09, 1.73, 2.51, 3.7, 4.53],
                    [0.0, -1.0, -9.0, -10.0, -13.0, -16.0, -21.6, -24.0]),
    'Bad_Urban': ([0.0, 0.1, 0.3, 0.5, 0.9, 1.3, 1.9, 2.6],
                  [0.0, -3.0, -5.0, -7.0, -9.0, -11.0, -13.0, -15.0]),
}
SLOT_SIZE = 14  # core/lte_receiver.py:233


class Numerology:
    """config.py:63-154 (LTEConfig._calculate_parameters)."""

    def __init__(self, bandwidth=5.0, delta_f=15.0, modulation='QPSK', cp_type='normal'):
        if modulation not in BITS_PER_SYMBOL:
            raise ValueError(f"Unsupported modulation: {modulation}")
        self.bandwidth, self.delta_f = bandwidth, delta_f
        self.modulation, self.cp_type = modulation, cp_type
        if bandwidth in LTE_PROFILES:                       # config.py:104-107
            self.Nc, self.N = LTE_PROFILES[bandwidth]
        else:                                               # config.py:110-111
            self.Nc = int((bandwidth * 1e3) / delta_f)
            self.N = int(2 ** np.ceil(np.log2(self.Nc)))
        self.fs = self.N * delta_f * 1e3                    # config.py:114
        if cp_type == 'extended':                           # config.py:136-145
            cp_us = CP_US['extended_15khz'] if delta_f == 15.0 else CP_US['extended_7.5khz']
        else:
            cp_us = CP_US['normal']
        self.cp_duration = cp_us
        self.cp_length = int(cp_us * 1e-6 * self.fs)        # config.py:124
        self.bits_per_symbol = BITS_PER_SYMBOL[modulation]
        self.L = self.N + self.cp_length


def grid_indices(N, Nc):
    """core/resource_mapper.py:45-74: bin classes in raw FFT-bin order."""
    gl = (N - Nc) // 2
    gr = N - Nc - gl
    dc = N // 2
    k = np.arange(N)
    useful = (k >= gl) & (k < N - gr) & (k != dc)
    pilot = useful & (((k - gl) % 6) == 3)
    data = useful & ~pilot
    return np.nonzero(data)[0], np.nonzero(pilot)[0]


def pilot_signs(cell_id, num_pilots):
    """core/resource_mapper.py:148-149: seed(cell_id); choice([1,-1], P)."""
    return np.random.RandomState(cell_id).choice([1, -1], size=num_pilots)


def pilots(cell_id, num_pilots):
    """core/resource_mapper.py:133,151."""
    return ((1 + 1j) / np.sqrt(2)) * pilot_signs(cell_id, num_pilots)


def constellation(modulation):
    """core/modulator.py:28-59: raster order, real outer loop / imag inner loop."""
    if modulation == 'QPSK':
        return np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
    if modulation == '16-QAM':
        lv, nrm = np.array([-3, -1, 1, 3]), np.sqrt(10)
    elif modulation == '64-QAM':
        lv, nrm = np.array([-7, -5, -3, -1, 1, 3, 5, 7]), np.sqrt(42)
    else:
        raise ValueError(f"unsupported modulation {modulation}")
    return (lv[:, None] + 1j * lv[None, :]).reshape(-1) / nrm


# ----------------------------------------------------------------------------
# L1: QAM map / demap                                (core/modulator.py:61-112)
# ----------------------------------------------------------------------------
def bits_to_indices(bits, b):
    """core/modulator.py:74-84: zero-pad to a multiple of b; MSB-first integer."""
    bits = np.asarray(bits).astype(np.int64)
    if len(bits) % b:
        bits = np.concatenate([bits, np.zeros(b - len(bits) % b, dtype=np.int64)])
    w = 1 << np.arange(b - 1, -1, -1)
    return (bits.reshape(-1, b) * w).sum(axis=1)


def indices_to_bits(idx, b):
    """core/modulator.py:109-110: format(idx, '0{b}b')."""
    idx = np.asarray(idx, dtype=np.int64)
    return ((idx[:, None] >> np.arange(b - 1, -1, -1)) & 1).reshape(-1)


def qam_map(bits, modulation):
    """core/modulator.py:61-88."""
    c = constellation(modulation)
    return c[bits_to_indices(bits, BITS_PER_SYMBOL[modulation]) % len(c)]


def qam_demap_indices(symbols, modulation):
    """core/modulator.py:103-106: argmin_i |c_i - y|, first minimum wins.

    Kept as a literal broadcast argmin (chunked) so that ties and the
    floating-point behaviour of np.abs() are the reference's.
    """
    c = constellation(modulation)
    symbols = np.asarray(symbols, dtype=complex)
    out = np.empty(len(symbols), dtype=np.int64)
    for s in range(0, len(symbols), 1 << 14):
        blk = symbols[s:s + (1 << 14)]
        out[s:s + len(blk)] = np.argmin(np.abs(c[None, :] - blk[:, None]), axis=1)
    return out


def qam_demap(symbols, modulation):
    """core/modulator.py:90-112."""
    return indices_to_bits(qam_demap_indices(symbols, modulation), BITS_PER_SYMBOL[modulation])


# ----------------------------------------------------------------------------
# SC-FDM M-point unitary DFT                    (core/dft_precoding.py:44-54,86-88)
# ----------------------------------------------------------------------------
def dft_precode(symbols, inverse=False):
    """Unitary M-point DFT of each row (last axis).  The reference multiplies by a
    dense matrix exp(-/+ j 2 pi k n / M)/sqrt(M) (core/dft_precoding.py:54,175);
    pocketfft computes the same sums more accurately."""
    M = symbols.shape[-1]
    if inverse:
        return np.fft.ifft(symbols, axis=-1) * np.sqrt(M)
    return np.fft.fft(symbols, axis=-1) / np.sqrt(M)


# ----------------------------------------------------------------------------
# L2: transmitter                      (core/modulator.py:214-302, resource_mapper.py:181-223)
# ----------------------------------------------------------------------------
def map_grid(data_symbols, num, cell_id=0):
    """ResourceMapper.map_symbols for a [S, Nd] block -> [S, N] grids."""
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    S = data_symbols.shape[0]
    grid = np.zeros((S, num.N), dtype=complex)
    grid[:, data_idx[:data_symbols.shape[1]]] = data_symbols
    grid[:, pilot_idx] = pilots(cell_id, len(pilot_idx))[None, :]
    return grid


def ofdm_modulate_grid(grid, num):
    """core/modulator.py:242-248: ifft * sqrt(N), prepend the last cp samples."""
    td = np.fft.ifft(grid, axis=-1) * np.sqrt(num.N)
    cp = num.cp_length
    if cp > 0:
        td = np.concatenate([td[..., -cp:], td], axis=-1)
    return td


def modulate_stream(bits, num, sc_fdm=False, mode='lte'):
    """OFDMModulator.modulate_stream (core/modulator.py:252-302).

    Returns (signal[S*L], qam_symbols[S, Nd]) for mode 'lte' (and SC-FDM);
    mode 'simple' puts the first Nc symbols on bins 0..Nc-1 (modulator.py:192-212).
    """
    b = num.bits_per_symbol
    bits = np.asarray(bits).astype(np.int64)
    if mode == 'simple':
        n_per = num.Nc
    else:
        data_idx, _ = grid_indices(num.N, num.Nc)
        n_per = len(data_idx)
    bits_per_ofdm = n_per * b
    S = int(np.ceil(len(bits) / bits_per_ofdm))
    if len(bits) < S * bits_per_ofdm:
        bits = np.concatenate([bits, np.zeros(S * bits_per_ofdm - len(bits), dtype=np.int64)])
    sym = qam_map(bits, num.modulation).reshape(S, n_per)
    if mode == 'simple':
        grid = np.zeros((S, num.N), dtype=complex)
        grid[:, :n_per] = sym
    else:
        pre = dft_precode(sym) if sc_fdm else sym
        grid = map_grid(pre, num)
    sig = ofdm_modulate_grid(grid, num)
    return sig.reshape(-1), sym


def papr(signal):
    """core/ofdm_core.py:131-140."""
    p = np.abs(signal) ** 2
    avg = np.mean(p)
    if avg > 0:
        lin = np.max(p) / avg
        return 10 * np.log10(lin), lin
    return 0.0, 1.0


def papr_per_symbol_no_cp(signal, num):
    """core/ofdm_system.py:173-229 (per-OFDM-symbol PAPR with the CP removed), linear ratio."""
    S = len(signal) // num.L
    x = signal[:S * num.L].reshape(S, num.L)[:, num.cp_length:]
    p = np.abs(x) ** 2
    return p.max(axis=1) / p.mean(axis=1)


def papr_per_symbol_db(signal, num, include_cp=False):
    """core/ofdm_system.py:116-171 (with the CP) and :173-229 (useful part only):
    -> (papr_db[S], peak[S], mean[S]); an all-zero symbol reports 0 dB (:212-217)."""
    S = len(signal) // num.L
    x = np.asarray(signal)[:S * num.L].reshape(S, num.L)
    if not include_cp:
        x = x[:, num.cp_length:]
    p = np.abs(x) ** 2
    peak, mean = p.max(axis=1), p.mean(axis=1)
    db = np.zeros(S)
    nz = mean > 0
    db[nz] = 10 * np.log10(peak[nz] / mean[nz])
    return db, peak, mean


def papr_histogram(papr_db, lo, step, bins):
    """Bin rule of lte_tx_papr / lte_histogram: floor((x-lo)/step) clamped into [0, bins-1]."""
    b = np.clip(np.floor((np.asarray(papr_db, dtype=np.float64) - lo) / step), 0, bins - 1).astype(np.int64)
    return np.bincount(b, minlength=bins).astype(np.int64)


def ccdf(papr_db, thresholds_db):
    """P(PAPR > threshold) -- what the reference's CCDF plots draw from the collected values."""
    v = np.sort(np.asarray(papr_db, dtype=np.float64))
    return 1.0 - np.searchsorted(v, thresholds_db, side='right') / max(len(v), 1)


# ----------------------------------------------------------------------------
# L2: channel             (core/rayleighchannel.py:20-58, core/channel.py:34-68,162-234)
# ----------------------------------------------------------------------------
def itu_taps(profile, fs, gain_conversions=2):
    """core/channel.py:172-186 then core/rayleighchannel.py:16 (and once more in
    core/channel.py:436 for spatial multiplexing): the dB table is converted to
    linear `gain_conversions` times.  Delays: core/rayleighchannel.py:52."""
    if profile not in ITU:
        raise ValueError(f"Perfil ITU no encontrado: {profile}")
    delays_us, power_db = ITU[profile]
    g = np.array(power_db, dtype=float)
    for _ in range(gain_conversions):
        g = 10 ** (g / 20)
    d = np.array([int(np.round(du * 1e-6 * fs)) for du in np.array(delays_us)])
    return d, g


def doppler_hz(frequency_ghz, velocity_kmh):
    """core/channel.py:118-143."""
    return ((velocity_kmh / 3.6) * (frequency_ghz * 1e9)) / 3e8


def jakes_fading(n_samples, fs, fD, phi):
    """core/rayleighchannel.py:27-41 with the 16 random phases `phi` injected."""
    n_s = len(phi)
    t = np.arange(n_samples) / fs
    alpha = 2 * np.pi * np.arange(1, n_s + 1) / n_s
    h = np.zeros(n_samples, dtype=complex)
    for n in range(n_s):
        h += np.exp(1j * (2 * np.pi * fD * np.cos(alpha[n]) * t + phi[n]))
    return h * np.sqrt(2 / n_s)


def rayleigh_filter(x, fs, fD, delays, gains, phases):
    """core/rayleighchannel.py:44-58; phases[tap][16] = 2*pi*rand(16) per tap."""
    n = len(x)
    y = np.zeros(n, dtype=complex)
    for i in range(len(delays)):
        fading = jakes_fading(n, fs, fD, phases[i])
        xd = np.concatenate([np.zeros(delays[i]), x])[:n]
        y += gains[i] * fading * xd
    return y


def awgn(x, snr_db, z_re, z_im):
    """core/channel.py:46-66 / :216-232 with the unit normals injected:
    noise_re = normal(0, sqrt(P/snr/2), L) == sqrt(P/snr/2) * z_re."""
    p = np.mean(np.abs(x) ** 2)
    sigma = np.sqrt((p / (10 ** (snr_db / 10))) / 2)
    return x + (sigma * z_re + 1j * (sigma * z_im))


class ReferenceDraws:
    """Reproduces the reference's draws from NumPy's legacy *global* RNG.

    Every ResourceMapper.map_symbols / LTEChannelEstimator.estimate_channel call
    re-seeds the global RNG with cell_id (core/resource_mapper.py:148), so after
    the transmitter has run the global state is always `seed(0); choice(P)`.
    Draw order per link afterwards (measured, see tests/golden/make_golden.py):
    taps x rand(16) (core/rayleighchannel.py:31), normal(L) for the real part,
    normal(L) for the imaginary part (core/channel.py:227-228).
    """

    def __init__(self, num_pilots, cell_id=0, global_seed=None):
        if global_seed is not None:
            # mode='simple' never calls map_symbols, so the global RNG is whatever the
            # caller left it at; the golden generator seeds it explicitly.
            self.rs = np.random.RandomState(global_seed)
        else:
            self.rs = np.random.RandomState(cell_id)
            self.rs.choice([1, -1], size=num_pilots)

    def phases(self, n_taps, n_s=16):
        return np.stack([2 * np.pi * self.rs.rand(n_s) for _ in range(n_taps)])

    def unit_normals(self, n):
        z_re = self.rs.standard_normal(n)
        z_im = self.rs.standard_normal(n)
        return z_re, z_im


# ----------------------------------------------------------------------------
# L2: receiver                               (core/lte_receiver.py:40-180,360-491)
# ----------------------------------------------------------------------------
def rx_fft_stream(signal, num):
    """core/lte_receiver.py:444-491: S = len//L (>=1), zero-pad, strip CP, fft/sqrt(N)."""
    L = num.L
    S = max(len(signal) // L, 1)
    buf = np.zeros(S * L, dtype=complex)
    m = min(len(signal), S * L)
    buf[:m] = signal[:m]
    x = buf.reshape(S, L)[:, num.cp_length:]
    return np.fft.fft(x, axis=-1) / np.sqrt(num.N)


def interpolate_channel(pilot_idx, h_p, N):
    """core/lte_receiver.py:98-133: edge hold + np.linspace between pilots."""
    h = np.zeros(N, dtype=complex)
    h[:pilot_idx[0]] = h_p[0]
    h[pilot_idx[-1]:] = h_p[-1]
    for i in range(len(pilot_idx) - 1):
        i1, i2 = pilot_idx[i], pilot_idx[i + 1]
        h[i1:i2 + 1] = np.linspace(h_p[i], h_p[i + 1], i2 - i1 + 1)
    return h


def estimate_channel(Y_symbol, num, cell_id=0):
    """core/lte_receiver.py:62-87: LS at pilots, then interpolation."""
    _, pilot_idx = grid_indices(num.N, num.Nc)
    h_p = Y_symbol[pilot_idx] / pilots(cell_id, len(pilot_idx))
    return interpolate_channel(pilot_idx, h_p, num.N)


def estimate_channel_periodic(Y, num, cell_id=0):
    """core/lte_receiver.py:360-411: estimate on symbol 14*j, hold for the slot."""
    S = Y.shape[0]
    H = np.zeros_like(Y)
    for s0 in range(0, S, SLOT_SIZE):
        H[s0:s0 + SLOT_SIZE] = estimate_channel(Y[s0], num, cell_id)[None, :]
    return H


def zf_equalize(Y, H, reg=1e-6):
    """core/lte_receiver.py:174."""
    return Y / (H + reg)


def mrc_combine(Y_data, H_data, reg=1e-10):
    """core/ofdm_core.py:1484-1532: sum_r conj(H_r) Y_r / (sum_r |H_r|^2 + reg).
    Y_data, H_data: [R, S, Nd].  Accumulation order over antennas is r = 0..R-1."""
    num_ = np.zeros(Y_data.shape[1:], dtype=complex)
    den = np.zeros(Y_data.shape[1:], dtype=float)
    for r in range(Y_data.shape[0]):
        num_ += np.conj(H_data[r]) * Y_data[r]
        den += np.abs(H_data[r]) ** 2
    return num_ / (den + reg)


def count_errors(bits_tx, bits_rx):
    """core/ofdm_core.py:712-718: pad/truncate rx to len(tx); sum(tx != rx)."""
    n = len(bits_tx)
    if len(bits_rx) < n:
        bits_rx = np.concatenate([bits_rx, np.zeros(n - len(bits_rx), dtype=bits_rx.dtype)])
    else:
        bits_rx = bits_rx[:n]
    return int(np.sum(np.asarray(bits_tx) != bits_rx)), bits_rx


# ----------------------------------------------------------------------------
# L3: end-to-end chains                 (core/ofdm_core.py:660-737, 1536-1679)
# ----------------------------------------------------------------------------
def channel_link(signal_tx, num, channel_type, snr_db, draws, itu_profile='Pedestrian_A',
                 frequency_ghz=2.0, velocity_kmh=0.0, gain_conversions=2, phases=None, z=None):
    """One ChannelSimulator.transmit (core/channel.py:334-345).  Draws come from
    `draws` (ReferenceDraws) unless phases / z=(z_re, z_im) are injected."""
    n = len(signal_tx)
    if channel_type == 'rayleigh_mp':
        d, g = itu_taps(itu_profile, num.fs, gain_conversions)
        fD = doppler_hz(frequency_ghz, velocity_kmh)
        ph = draws.phases(len(d)) if phases is None else phases
        faded = rayleigh_filter(signal_tx, num.fs, fD, d, g, ph)
    else:   # unknown channel types mean AWGN (core/ofdm_core.py:644-654)
        faded = signal_tx
    z_re, z_im = draws.unit_normals(n) if z is None else z
    return awgn(faded, snr_db, z_re, z_im), faded


def simulate_siso(bits, snr_db, num, channel_type='awgn', mode='lte', sc_fdm=False,
                  equalize=True, itu_profile='Pedestrian_A', frequency_ghz=2.0,
                  velocity_kmh=0.0, phases=None, z=None, draws=None):
    """OFDMSimulator.simulate_siso (core/ofdm_core.py:660-737) with the LTE receiver
    of core/lte_receiver.py:235-358 and core/demodulator.py:138-147."""
    bits = np.asarray(bits).astype(np.int64)
    if bits.size == 0:
        raise ValueError("Bits array cannot be empty")
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    signal_tx, sym_tx = modulate_stream(bits, num, sc_fdm=sc_fdm, mode=mode)
    papr_db, papr_lin = papr(signal_tx)
    if draws is None and (phases is None or z is None):
        draws = ReferenceDraws(len(pilot_idx))
    signal_rx, faded = channel_link(signal_tx, num, channel_type, snr_db, draws, itu_profile,
                                    frequency_ghz, velocity_kmh, phases=phases, z=z)
    Y = rx_fft_stream(signal_rx, num)
    if mode == 'simple':                       # core/demodulator.py:114-118,149-184
        sym_rx = Y[:, :num.Nc].reshape(-1)
        H = None
    else:
        H = estimate_channel_periodic(Y, num)
        Yeq = zf_equalize(Y, H) if equalize else Y
        sym_rx = Yeq[:, data_idx]
        if sc_fdm:                             # core/lte_receiver.py:319-333
            sym_rx = dft_precode(sym_rx, inverse=True)
        sym_rx = sym_rx.reshape(-1)
    bits_rx_all = qam_demap(sym_rx, num.modulation)
    errors, bits_rx = count_errors(bits, bits_rx_all)
    return dict(signal_tx=signal_tx, symbols_tx=sym_tx, signal_faded=faded, signal_rx=signal_rx,
                Y=Y, H=H, symbols_rx=sym_rx, bits_rx=bits_rx, errors=errors,
                ber=errors / len(bits), papr_db=papr_db, papr_linear=papr_lin)


def simulate_simo(bits, snr_db, num, num_rx=2, channel_type='awgn', itu_profile='Pedestrian_A',
                  frequency_ghz=2.0, velocity_kmh=0.0, phases=None, z=None):
    """OFDMSimulator.simulate_simo (core/ofdm_core.py:1536-1679), sequential branch.
    phases: [R][taps][16]; z: [R][2][n] unit normals (re then im), or None to
    reproduce the reference's own draws."""
    bits = np.asarray(bits).astype(np.int64)
    if bits.size == 0:
        raise ValueError("Bits array cannot be empty")
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    signal_tx, sym_tx = modulate_stream(bits, num)
    papr_db, papr_lin = papr(signal_tx)
    draws = ReferenceDraws(len(pilot_idx)) if (phases is None or z is None) else None
    rx_list, faded_list = [], []
    for r in range(num_rx):                    # core/ofdm_core.py:396-410
        rx, faded = channel_link(signal_tx, num, channel_type, snr_db, draws, itu_profile,
                                 frequency_ghz, velocity_kmh,
                                 phases=None if phases is None else phases[r],
                                 z=None if z is None else (z[r][0], z[r][1]))
        rx_list.append(rx)
        faded_list.append(faded)
    Y = np.stack([rx_fft_stream(rx, num) for rx in rx_list])                 # [R,S,N]
    H = np.stack([estimate_channel_periodic(Y[r], num) for r in range(num_rx)])
    comb = mrc_combine(Y[:, :, data_idx], H[:, :, data_idx]).reshape(-1)
    bits_rx_all = qam_demap(comb, num.modulation)
    errors, bits_rx = count_errors(bits, bits_rx_all)
    return dict(signal_tx=signal_tx, symbols_tx=sym_tx, signal_faded=np.stack(faded_list),
                signal_rx=np.stack(rx_list), Y=Y, H=H, symbols_combined=comb,
                bits_rx=bits_rx, errors=errors, ber=errors / len(bits),
                papr_db=papr_db, papr_linear=papr_lin)


# ----------------------------------------------------------------------------
# SFBC-Alamouti transmit diversity      (core/sfbc_alamouti.py, core/ofdm_core.py:434-543,1850-2258)
# ----------------------------------------------------------------------------
def sfbc_encode(symbols):
    """core/sfbc_alamouti.py:45-78: pairs (k, k+1): TX0 [s0, -s1*], TX1 [s1, s0*]."""
    symbols = np.asarray(symbols, dtype=complex)
    if symbols.shape[-1] % 2:
        raise ValueError(f"Number of symbols must be even for Alamouti coding, got {symbols.shape[-1]}")
    s0, s1 = symbols[..., 0::2], symbols[..., 1::2]
    tx0 = np.empty_like(symbols)
    tx1 = np.empty_like(symbols)
    tx0[..., 0::2], tx0[..., 1::2] = s0, -np.conj(s1)
    tx1[..., 0::2], tx1[..., 1::2] = s1, np.conj(s0)
    return tx0, tx1


def sfbc_decode(rx, H0, H1, reg=1e-10):
    """core/sfbc_alamouti.py:80-163."""
    rx, H0, H1 = (np.asarray(a, dtype=complex) for a in (rx, H0, H1))
    if rx.shape[-1] % 2:
        raise ValueError(f"Number of RX symbols must be even, got {rx.shape[-1]}")
    ra, rb = rx[..., 0::2], rx[..., 1::2]
    h0a, h0b, h1a, h1b = H0[..., 0::2], H0[..., 1::2], H1[..., 0::2], H1[..., 1::2]
    norm = np.abs((h0a + h0b) / 2) ** 2 + np.abs((h1a + h1b) / 2) ** 2 + reg
    out = np.empty_like(rx)
    out[..., 0::2] = (np.conj(h0a) * ra + h1b * np.conj(rb)) / norm
    out[..., 1::2] = (np.conj(h1a) * ra - h0b * np.conj(rb)) / norm
    return out


def mimo_pilot_layout(num_tx, pilot_idx):
    """core/mimo_channel_estimator_periodic.py:75-106 and core/sfbc_alamouti.py:244-262: TX t owns
    pilot_idx[t % step :: step], step = min(num_tx, 4), with the sign stream of cell t % 4."""
    step = num_tx if num_tx <= 4 else 4
    own = [pilot_idx[t % step::step] for t in range(num_tx)]
    vals = [pilots(t % 4, len(own[t])) for t in range(num_tx)]
    return own, vals


def mimo_estimate_from_grid(Y, num_tx, num):
    """core/mimo_channel_estimator_periodic.py:108-185: Y [R, N] -> H [R, T, N]."""
    _, pilot_idx = grid_indices(num.N, num.Nc)
    own, vals = mimo_pilot_layout(num_tx, pilot_idx)
    Y = np.atleast_2d(Y)
    H = np.zeros((Y.shape[0], num_tx, num.N), dtype=complex)
    for r in range(Y.shape[0]):
        for t in range(num_tx):
            H[r, t] = interpolate_channel(own[t], Y[r, own[t]] / vals[t], num.N)
    return H


class SfbcDraws:
    """Draw order of OFDMChannel.transmit_mimo (core/ofdm_core.py:470-541) after the SFBC
    transmitter left the global RNG at seed(1); choice(len(pilot_idx[1::2]))
    (core/sfbc_alamouti.py:252-256)."""

    def __init__(self, num_pilots):
        self.rs = np.random.RandomState(1)
        self.rs.choice([1, -1], size=len(np.arange(num_pilots)[1::2]))

    def phases(self, n_taps, n_s=16):
        return np.stack([2 * np.pi * self.rs.rand(n_s) for _ in range(n_taps)])

    def unit_normals(self, n):
        return self.rs.standard_normal(n), self.rs.standard_normal(n)


def simulate_sfbc(bits, snr_db, num, num_rx=1, channel_type='awgn', itu_profile='Pedestrian_A',
                  frequency_ghz=2.0, velocity_kmh=0.0, draws=None):
    """OFDMSimulator.simulate_miso / simulate_mimo (core/ofdm_core.py:1850-2258) with the periodic
    estimator restated as intended (the HEAD version unpacks 3 values from a 2-value return,
    core/mimo_channel_estimator_periodic.py:219-222 vs :185 -- SURVEY 0.9): estimate on symbol 14j
    with estimate_channel_from_grid, hold for the slot."""
    bits = np.asarray(bits).astype(np.int64)
    if bits.size == 0:
        raise ValueError("Bits array cannot be empty")
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    nd2 = len(data_idx) - (len(data_idx) % 2)                   # core/sfbc_alamouti.py:196-200
    didx = data_idx[:nd2]
    b = num.bits_per_symbol
    S = int(np.ceil(len(bits) / (nd2 * b)))
    padded = np.concatenate([bits, np.zeros(S * nd2 * b - len(bits), dtype=np.int64)])
    sym = qam_map(padded, num.modulation).reshape(S, nd2)
    tx0d, tx1d = sfbc_encode(sym)
    own, vals = mimo_pilot_layout(2, pilot_idx)
    grids = np.zeros((2, S, num.N), dtype=complex)
    grids[0][:, didx], grids[1][:, didx] = tx0d, tx1d
    grids[0][:, own[0]], grids[1][:, own[1]] = vals[0], vals[1]
    sig = ofdm_modulate_grid(grids, num)                        # [2, S, L]
    papr_db = [float(np.mean([papr(sig[t, s])[0] for s in range(S)])) for t in range(2)]
    tx = sig.reshape(2, -1)
    n = tx.shape[1]
    if draws is None:
        draws = SfbcDraws(len(pilot_idx))
    snr_lin = 10 ** (snr_db / 10)
    rx_list, links = [], np.zeros((num_rx, 2, n), dtype=complex)
    chan_mat = np.zeros((num_rx, 2), dtype=complex)
    for r in range(num_rx):                                     # core/ofdm_core.py:470-541
        acc = np.zeros(n, dtype=complex)
        for t in range(2):
            if channel_type == 'rayleigh_mp':
                d, g = itu_taps(itu_profile, num.fs)
                fD = doppler_hz(frequency_ghz, velocity_kmh)
                faded = rayleigh_filter(tx[t], num.fs, fD, d, g, draws.phases(len(d)))
                z_re, z_im = draws.unit_normals(n)
                link = awgn(faded, 100.0, z_re, z_im)           # the per-link 100 dB noise
                tp, rp = np.mean(np.abs(tx[t]) ** 2), np.mean(np.abs(link) ** 2)
                chan_mat[r, t] = np.sqrt(rp / tp) * np.exp(1j * np.angle(np.mean(link * np.conj(tx[t]))))
            else:
                h = 1.0 + 0j if t == 0 else np.exp(1j * (t * np.pi / 2))
                link = tx[t] * h
                chan_mat[r, t] = h
            links[r, t] = link
            acc += link
        p = np.mean(np.abs(acc) ** 2)
        sigma = np.sqrt(((p / 2) / snr_lin) / 2)
        z_re, z_im = draws.unit_normals(n)
        rx_list.append(acc + (sigma * z_re + 1j * (sigma * z_im)))
    rx = np.stack(rx_list)
    Y = np.stack([rx_fft_stream(rx[r], num) for r in range(num_rx)])         # [R, S, N]
    Srx = Y.shape[1]
    H = np.zeros((num_rx, 2, Srx, num.N), dtype=complex)
    for r in range(num_rx):
        for s0 in range(0, Srx, SLOT_SIZE):
            H[r, :, s0:s0 + SLOT_SIZE] = mimo_estimate_from_grid(Y[r, s0], 2, num)[0][:, None, :]
    dec = np.stack([sfbc_decode(Y[r][:, didx], H[r, 0][:, didx], H[r, 1][:, didx]) for r in range(num_rx)])
    comb = np.mean(dec, axis=0).reshape(-1)                     # core/ofdm_core.py:2204 (plain average)
    bits_rx_all = qam_demap(comb, num.modulation)
    errors, bits_rx = count_errors(bits, bits_rx_all)
    return dict(signal_tx=tx, signal_rx=rx, links=links, Y=Y, H=H, symbols=comb, bits_rx=bits_rx, errors=errors,
                ber=errors / len(bits), channel_matrix=chan_mat, papr_db_tx0=papr_db[0], papr_db_tx1=papr_db[1])


# ----------------------------------------------------------------------------
# Spatial multiplexing (TM4-like)   (core/ofdm_core.py:2489-2815, core/mimo_detector.py,
#                                    core/layer_mapper.py, core/codebook_lte.py, core/rank_adaptation.py)
# ----------------------------------------------------------------------------
def _ph(theta):
    return np.exp(1j * theta)


def codebook(num_tx, rank):
    """core/codebook_lte.py:59-293 ('TM4'; rank 1 re-uses the TM6 vectors, :115-118)."""
    if rank < 1 or rank > min(num_tx, 4):
        raise ValueError(f"TM4 con {num_tx} antenas soporta rank 1-{min(num_tx, 4)}, recibido rank={rank}")
    cb = []
    if rank == 1:
        if num_tx == 2:
            cb = [np.array([[1], [v]]) / np.sqrt(2) for v in (1, -1, 1j, -1j)]
        elif num_tx in (4, 8):
            nrm = 2 if num_tx == 4 else np.sqrt(8)
            cb = [_ph(2 * np.pi * i * np.arange(num_tx) / 16).reshape(-1, 1) / nrm for i in range(16)]
    elif rank == 2:
        if num_tx == 2:
            cb = [np.array([[1, 0], [0, 1]]), np.array([[1, 1], [1, -1]]) / np.sqrt(2),
                  np.array([[1, 1], [1j, -1j]]) / np.sqrt(2)]
        elif num_tx == 4:
            e = [_ph(2 * np.pi * i / 4) for i in range(4)]
            cb += [np.array([[1, 0], [p, 0], [0, 1], [0, p]]) / np.sqrt(2) for p in e]
            cb += [np.array([[1, 1], [p, -p], [1, -1], [p, p]]) / 2 for p in e]
            cb += [np.array([[1, 0], [0, 1], [p, 0], [0, p]]) / np.sqrt(2) for p in e]
            cb += [np.array([[1, 1], [1, -1], [p, p], [p, -p]]) / 2 for p in e]
        elif num_tx == 8:
            for i in range(16):
                W = np.zeros((8, 2), dtype=complex)
                W[0:4, 0] = W[4:8, 1] = _ph(2 * np.pi * i / 16 * np.arange(4)) / np.sqrt(4)
                cb.append(W)
    elif rank == 3:
        if num_tx == 4:
            for i in range(8):
                p = _ph(2 * np.pi * i / 8)
                cb.append(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [p, p, p]]) / np.sqrt(2))
        elif num_tx == 8:
            for i in range(16):
                th = 2 * np.pi * i / 16
                v = np.array([1, _ph(th), _ph(2 * th)]) / np.sqrt(3)
                W = np.zeros((8, 3), dtype=complex)
                W[0:3, 0] = v
                W[3:6, 1] = v
                W[5:8, 2] = v
                cb.append(W)
    elif rank == 4:
        if num_tx == 4:
            ij = np.outer(np.arange(4), np.arange(4))
            cb = [np.eye(4, dtype=complex), np.exp(-2j * np.pi * ij / 4) / 2,
                  np.array([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]) / 2,
                  np.array([[1, 1, 1, 1], [1, 1j, -1, -1j], [1, -1, 1, -1], [1, -1j, -1, 1j]]) / 2]
        elif num_tx == 8:
            for i in range(8):
                th = 2 * np.pi * i / 8
                W = np.zeros((8, 4), dtype=complex)
                for l in range(4):
                    W[2 * l:2 * l + 2, l] = np.array([1, _ph(th * (l + 1))]) / np.sqrt(2)
                cb.append(W)
    if not cb:
        raise ValueError(f"num_tx={num_tx} no soportado en TM4 Rank-{rank}")
    return cb


def rank_feedback(H, num_tx, num_rx, snr_db, rank_threshold=0.15):
    """RankAdaptation.get_feedback (core/rank_adaptation.py:69-265): RI from the eigenvalues of
    H^H H with SNR gating, PMI maximising log2 det(I + snr/rank H_eff H_eff^H)."""
    max_rank = min(num_tx, num_rx, 4)
    ev = np.sort(np.linalg.eigvalsh(H.conj().T @ H))[::-1]
    if ev[0] < 1e-10:
        ri = 1
    else:
        ri = min(int(np.sum(ev / ev[0] > rank_threshold)), max_rank)
        if snr_db < 5:
            ri = 1
        elif snr_db < 10:
            ri = min(ri, 2)
        ri = max(1, ri)
    snr_lin = 10 ** (snr_db / 10)
    best, best_val = 0, -np.inf
    cb = codebook(num_tx, ri)
    for pmi, W in enumerate(cb):
        He = H @ W
        val = np.log2(np.linalg.det(np.eye(num_rx) + (snr_lin / ri) * (He @ He.conj().T)))
        if val > best_val:
            best_val, best = val, pmi
    return ri, best, cb[best]


def layer_map(symbols, rank):
    """core/layer_mapper.py:35-86: round robin, zero padded -> [rank, ceil(n / rank)]."""
    symbols = np.asarray(symbols)
    if rank == 1:
        return symbols.reshape(1, -1)
    if len(symbols) % rank:
        symbols = np.concatenate([symbols, np.zeros(rank - len(symbols) % rank, dtype=symbols.dtype)])
    return symbols.reshape(-1, rank).T


def layer_demap(layers, original_length=None):
    """core/layer_mapper.py:88-115."""
    s = np.asarray(layers).T.flatten()
    return s if original_length is None else s[:original_length]


def mimo_detect(y, H_eff, sigma2, detector, constellation_points=None):
    """MIMODetector._detect_single for one subcarrier (core/mimo_detector.py:99-369).
    y [R]; H_eff [R, L]."""
    L = H_eff.shape[1]
    det = detector.upper()
    if det in ('MMSE', 'IRC'):
        return np.linalg.inv(H_eff.conj().T @ H_eff + sigma2 * np.eye(L)) @ H_eff.conj().T @ y
    if det == 'ZF':
        return np.linalg.pinv(H_eff) @ y
    if det == 'MRC':
        if L != 1:
            raise ValueError("MRC solo soporta num_layers=1 (rank-1)")
        h = H_eff[:, 0]
        return np.array([np.dot(h.conj() / (np.linalg.norm(h) ** 2), y)])
    if det == 'SIC':
        norms = np.array([np.linalg.norm(H_eff[:, i]) ** 2 for i in range(L)])
        sinr = np.array([norms[i] / (norms.sum() - norms[i] + sigma2 + 1e-10) for i in range(L)])
        order = np.argsort(sinr)[::-1]
        y_res, H_rem, remaining = y.copy(), H_eff.copy(), list(range(L))
        s_hat = np.zeros(L, dtype=complex)
        for it in range(L):
            layer = order[it]
            rel = remaining.index(layer)
            if H_rem.shape[1] == 1:
                h = H_rem[:, 0]
                s = np.vdot(h, y_res) / (np.linalg.norm(h) ** 2 + sigma2)
            else:
                s = (np.linalg.inv(H_rem.conj().T @ H_rem + sigma2 * np.eye(H_rem.shape[1])) @ H_rem.conj().T @ y_res)[rel]
            s_hard = constellation_points[np.argmin(np.abs(constellation_points - s))]
            s_hat[layer] = s_hard
            y_res = y_res - H_eff[:, layer] * s_hard
            if it < L - 1:
                H_rem = np.delete(H_rem, rel, axis=1)
                remaining.pop(rel)
        return s_hat
    raise ValueError(f"Detector '{detector}' no soportado")


class SmDraws:
    """Global-RNG draw order of simulate_spatial_multiplexing: H_initial = randn(R,T), randn(R,T)
    from the caller's state (core/ofdm_core.py:2574), then the transmitter re-seeds with
    cell (T-1) % 4 (:2651-2654), then transmit_spatial_multiplexing draws (core/channel.py:441-491)."""

    def __init__(self, num_tx, num_pilots, global_seed):
        self.h_rs = np.random.RandomState(global_seed)
        t = num_tx - 1
        step = num_tx if num_tx <= 4 else 4
        self.rs = np.random.RandomState(t % 4)
        self.rs.choice([1, -1], size=len(np.arange(num_pilots)[t % step::step]))

    def h_initial(self, num_rx, num_tx):
        return (self.h_rs.randn(num_rx, num_tx) + 1j * self.h_rs.randn(num_rx, num_tx)) / np.sqrt(2 * num_tx)


def simulate_sm(bits, num, num_tx=4, num_rx=2, rank='adaptive', detector='MMSE', snr_db=15.0,
                channel_type='awgn', itu_profile='Pedestrian_A', velocity_kmh=3.0, frequency_ghz=2.0,
                enable_csi_feedback=True, global_seed=0):
    """simulate_spatial_multiplexing (core/ofdm_core.py:2489-2815)."""
    bits = np.asarray(bits).astype(np.int64)
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    Nd, b = len(data_idx), num.bits_per_symbol
    S = int(np.ceil(len(bits) / (Nd * b)))
    padded = np.concatenate([bits, np.zeros(S * Nd * b - len(bits), dtype=np.int64)])
    draws = SmDraws(num_tx, len(pilot_idx), global_seed)
    H_init = draws.h_initial(num_rx, num_tx)
    if rank == 'adaptive' and enable_csi_feedback:
        ri, pmi, W = rank_feedback(H_init, num_tx, num_rx, snr_db)
    else:
        ri = int(rank) if rank != 'adaptive' else min(num_tx, num_rx)
        pmi, W = 0, codebook(num_tx, ri)[0]
    W = np.asarray(W, dtype=complex)
    sym = qam_map(padded, num.modulation).reshape(S, Nd)
    own, vals = mimo_pilot_layout(num_tx, pilot_idx)
    npos = int(np.ceil(Nd / ri))
    grids = np.zeros((num_tx, S, num.N), dtype=complex)
    for s in range(S):
        layers = layer_map(sym[s], ri)                            # [ri, npos]
        grids[:, s, data_idx[:npos]] = W @ layers                 # core/ofdm_core.py:2630-2640
        for t in range(num_tx):
            grids[t, s, own[t]] = vals[t]
    tx = ofdm_modulate_grid(grids, num).reshape(num_tx, -1)
    n = tx.shape[1]
    rs = draws.rs
    rx = np.zeros((num_rx, n), dtype=complex)
    Hc = np.zeros((num_rx, num_tx), dtype=complex)
    if channel_type == 'rayleigh_mp':
        d, g = itu_taps(itu_profile, num.fs, gain_conversions=3)  # core/channel.py:436,444
        fD = doppler_hz(frequency_ghz, velocity_kmh)
        for r in range(num_rx):
            for t in range(num_tx):
                ph = np.stack([2 * np.pi * rs.rand(16) for _ in range(len(d))])
                rx[r] += rayleigh_filter(tx[t], num.fs, fD, d, g, ph)
                ph2 = np.stack([2 * np.pi * rs.rand(16) for _ in range(len(d))])    # impulse_response(N=1)
                Hc[r, t] = g[0] * jakes_fading(1, num.fs, fD, ph2[0])[0]
    else:
        for r in range(num_rx):
            for t in range(num_tx):
                h = rs.normal(0, 1 / np.sqrt(2)) + 1j * rs.normal(0, 1 / np.sqrt(2))
                Hc[r, t] = h
                rx[r] += h * tx[t]
    snr_lin = 10 ** (snr_db / 10)
    for r in range(num_rx):
        p = np.mean(np.abs(rx[r]) ** 2)
        sg = np.sqrt((p / snr_lin) / 2)
        rx[r] = rx[r] + (sg * rs.standard_normal(n) + 1j * (sg * rs.standard_normal(n)))
    Y = np.stack([rx_fft_stream(rx[r], num)[:S] for r in range(num_rx)])       # [R, S, N]
    sigma2 = 10 ** (-snr_db / 10)
    cpts = constellation(num.modulation)
    out = np.zeros((S, Nd), dtype=complex)
    H_all = np.zeros((S, num_rx, num_tx, num.N), dtype=complex)
    for s in range(min(S, Y.shape[1])):
        H = mimo_estimate_from_grid(Y[:, s], num_tx, num)         # every OFDM symbol (:2743-2752)
        H_all[s] = H
        lay = np.zeros((ri, Nd), dtype=complex)
        for k in range(npos):                                     # bins past npos are discarded by the demap
            lay[:, k] = mimo_detect(Y[:, s, data_idx[k]], H[:, :, data_idx[k]] @ W, sigma2, detector, cpts)
        out[s] = layer_demap(lay[:, :npos], Nd)
    symbols = out.reshape(-1)
    bits_rx_all = qam_demap(symbols, num.modulation)
    errors, bits_rx = count_errors(bits, bits_rx_all)
    return dict(signal_tx=tx, signal_rx=rx, Y=Y, H=H_all, symbols=symbols, bits_rx=bits_rx, errors=errors,
                ber=errors / len(bits), rank=ri, pmi=pmi, W=W, channel_matrix=Hc)


# ----------------------------------------------------------------------------
# Beamforming (TM6-like rank-1 precoding over a flat channel)
#   (core/ofdm_core.py:2260-2477, core/csi_feedback.py, core/beamforming_precoder.py)
# ----------------------------------------------------------------------------
def select_best_pmi(H, cb):
    """LTECodebook.select_best_pmi, metric 'capacity' (core/codebook_lte.py:332-373):
    argmax_i ||H w_i||^2, first maximum wins (strict >)."""
    best, best_m = 0, -np.inf
    for i, W in enumerate(cb):
        m = np.sum(np.abs(H @ W) ** 2)
        if m > best_m:
            best_m, best = m, i
    return best, best_m


def mrt_weights(H):
    """BeamformingPrecoder.calculate_mrt_weights (core/beamforming_precoder.py:41-66): conjugate of
    the RX-averaged channel row, unit norm."""
    h = np.conj(np.mean(H, axis=0)) if H.ndim == 2 else np.conj(H)
    return (h / np.sqrt(np.sum(np.abs(h) ** 2))).reshape(-1, 1)


def beamforming_gain_db(H, W, num_tx):
    """calculate_beamforming_gain (core/beamforming_precoder.py:176-201)."""
    return 10 * np.log10(np.sum(np.abs(H @ W) ** 2) / (np.sum(np.abs(H) ** 2) / num_tx))


def sinr_to_cqi(sinr_db):
    """CSIFeedback._sinr_to_cqi (core/csi_feedback.py:106-137): 2 dB steps from -6 dB, CQI 0..15."""
    if sinr_db < -6.0:
        return 0
    return int(min(15, np.floor((sinr_db + 6.0) / 2.0) + 1))


def beamforming_update_period(velocity_kmh, frequency_ghz=2.0):
    """AdaptiveBeamforming._calculate_update_period (core/beamforming_precoder.py:231-263)."""
    fd = (velocity_kmh / 3.6) * (frequency_ghz * 1e9) / 3e8
    if fd == 0:
        return 100
    return int(np.clip(int(0.1 * (9 / (16 * np.pi * fd)) / (1 / 15000)), 1, 140))


def simulate_beamforming(bits, snr_db, num, num_tx=2, num_rx=1, update_mode='adaptive', global_seed=0):
    """OFDMSimulator.simulate_beamforming (core/ofdm_core.py:2260-2477).  Draw order from the caller's
    global RNG state (nothing on this path re-seeds it): randn(R,T), randn(R,T) for H, then per OFDM
    symbol randn(R,Nd), randn(R,Nd) for the noise.  The codebook is the rank-1 one for both 'TM6'
    and 'TM4' (core/codebook_lte.py:114-118)."""
    rs = np.random.RandomState(global_seed)
    bits = np.asarray(bits).astype(np.int64)
    n = len(bits)
    data_idx, _ = grid_indices(num.N, num.Nc)
    Nd, b = len(data_idx), num.bits_per_symbol
    S = int(np.ceil(n / (Nd * b)))
    padded = np.concatenate([bits, np.zeros(S * Nd * b - n, dtype=np.int64)])
    sym = qam_map(padded, num.modulation).reshape(S, Nd)
    H = (rs.randn(num_rx, num_tx) + 1j * rs.randn(num_rx, num_tx)) / np.sqrt(2)
    cb = codebook(num_tx, 1)
    nv = 10 ** (-snr_db / 10)
    rx_all, gains, pmis = [], [], []
    W = None
    for s in range(S):
        pmi, _ = select_best_pmi(H, cb)                       # CSI feedback every symbol (:2366-2369)
        pmis.append(pmi)
        if update_mode == 'adaptive':
            W = mrt_weights(H)                                # update_precoder every symbol (:2372-2375)
            gains.append(beamforming_gain_db(H, W, num_tx))
        else:
            W = cb[pmi]
            gains.append(0.0)                                 # precoder.W stays None => 0.0 (:190-191)
        x = W @ sym[s].reshape(1, -1)                         # [T, Nd]
        rx = np.zeros((num_rx, Nd), dtype=complex)
        for r in range(num_rx):
            for t in range(num_tx):
                rx[r] += H[r, t] * x[t]
        noise = (rs.randn(num_rx, Nd) + 1j * rs.randn(num_rx, Nd)) * np.sqrt(nv / 2)
        rx_all.append(rx + noise)
    H_eff = H @ W                                             # last symbol's precoder (:2411)
    pn = np.sum(np.abs(H_eff) ** 2)
    eq = []
    for s in range(S):
        c = np.zeros(Nd, dtype=complex)
        for r in range(num_rx):
            c += np.conj(H_eff[r, 0]) * rx_all[s][r]
        eq.append(c / pn)
    symbols = np.concatenate(eq)
    bits_rx = qam_demap(symbols, num.modulation)[:n]
    errors = int(np.sum(bits != bits_rx))
    return dict(symbols=symbols, bits_rx=bits_rx, errors=errors, ber=errors / n, channel_matrix=H, W=W,
                H_eff=H_eff, pmi_history=pmis, beamforming_gain_db=float(np.mean(gains)),
                unique_pmis=len(set(pmis)))


# ----------------------------------------------------------------------------
# Coded chain: CRC + segmentation + turbo code + rate matching + soft demapping
#   (core/ofdm_core.py:739-1338, core/channel_coding/*.py)
# ----------------------------------------------------------------------------
CRC24A_POLY, CRC24B_POLY = 0x1864CFB, 0x1800063      # core/channel_coding/crc.py:32-33


def crc24(bits, poly=CRC24A_POLY):
    """_calculate_crc (core/channel_coding/crc.py:78-124): remainder of bits * D^24 by the generator,
    zero initial state, MSB first -- as a 24-bit shift register."""
    reg = 0
    for b in np.asarray(bits).astype(np.int64):
        top = ((reg >> 23) & 1) ^ int(b)
        reg = (reg << 1) & 0xFFFFFF
        if top:
            reg ^= poly & 0xFFFFFF
    return np.array([(reg >> (23 - i)) & 1 for i in range(24)], dtype=np.uint8)


TURBO_K = ([40 + 8 * i for i in range(60)] + [528 + 16 * i for i in range(32)] +
           [1056 + 32 * i for i in range(32)] + [2112 + 64 * i for i in range(64)])   # segmentation.py:28-45

def find_interleaver_size(n):
    for k in TURBO_K:
        if k >= n:
            return k
    raise ValueError(f"No valid interleaver size found for min_size={n}")


def segmentation_layout(B):
    """segment_code_blocks (core/channel_coding/segmentation.py:66-199) as a layout:
    list of (K_r, F_r filler bits, info bits, has_crc24b) per code block."""
    Z, L = 6144, 24
    if B <= Z:
        K = find_interleaver_size(B)
        return [(K, K - B, B, False)]
    C = int(np.ceil(B / (Z - L)))
    Bp = B + C * L
    Kp = find_interleaver_size(int(np.ceil(Bp / C)))
    i = TURBO_K.index(Kp) - 1
    Km = TURBO_K[i] if i >= 0 else Kp
    dK = Kp - Km
    Cm = (C * Kp - Bp) // dK if dK > 0 else 0
    out, remaining = [], B
    for r in range(C):
        Kr = Km if r < Cm else Kp
        n = remaining if r == C - 1 else min(Kr - L, remaining // (C - r))
        remaining -= n
        out.append((Kr, (Kr - L) - n, n, True))
    return out


def segment_code_blocks(tb):
    blocks, pos = [], 0
    for Kr, F, n, has_crc in segmentation_layout(len(tb)):
        body = np.zeros(Kr - (24 if has_crc else 0), dtype=np.uint8)
        body[F:F + n] = tb[pos:pos + n]
        pos += n
        blocks.append(np.concatenate([body, crc24(body, CRC24B_POLY)]) if has_crc else body)
    return blocks


def desegment_code_blocks(blocks, B):
    out = []
    for cb, (Kr, F, n, has_crc) in zip(blocks, segmentation_layout(B)):
        out.append(cb[F:F + n])
    return np.concatenate(out)


def qpp_indices(K, f1, f2):
    i = np.arange(K, dtype=np.int64)
    return (f1 * i + f2 * i * i) % K


def rsc_encode(u):
    """rsc_encode (core/channel_coding/turbo_encoder.py:112-169): the 'systematic' output is the
    FEEDBACK bit a_k = u_k + s1 + s2 (not u_k); parity = a_k + s0 + s2; 3 termination steps."""
    s0 = s1 = s2 = 0
    sys_, par = [], []
    for bit in list(np.asarray(u).astype(int)) + [None] * 3:
        if bit is None:
            bit = (s1 + s2) % 2
        fb = (bit + s1 + s2) % 2
        sys_.append(fb)
        par.append((fb + s0 + s2) % 2)
        s0, s1, s2 = fb, s0, s1
    return np.array(sys_, dtype=np.uint8), np.array(par, dtype=np.uint8)


def turbo_encode(u, f1, f2):
    """turbo_encode (turbo_encoder.py:172-259): [s_k, p1_k, p2_k] interleaved for k < K, then
    tail_sys1(3) tail_par1(3) tail_sys2(3) tail_par2(3)."""
    K = len(u)
    s1_, p1 = rsc_encode(u)
    s2_, p2 = rsc_encode(np.asarray(u)[qpp_indices(K, f1, f2)])
    out = np.zeros(3 * K + 12, dtype=np.uint8)
    out[0:3 * K:3], out[1:3 * K:3], out[2:3 * K:3] = s1_[:K], p1[:K], p2[:K]
    out[3 * K:] = np.concatenate([s1_[K:], p1[K:], s2_[K:], p2[K:]])
    return out


_SBI_P = np.array([0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                   1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31])


def sub_block_permutation(n):
    """sub_block_interleaver (core/channel_coding/rate_matching.py:27-77) as an index vector:
    out[j] = in[perm[j]].  The matrix is filled COLUMN by column with the nulls at the end,
    columns are permuted, rows are read out and nulls dropped (the reference's own variant)."""
    R = int(np.ceil(n / 32))
    m = np.full(R * 32, -1, dtype=np.int64)
    m[:n] = np.arange(n)
    m = m.reshape(32, R).T[:, _SBI_P].reshape(-1)
    return m[m >= 0]


def rate_match_table(K):
    """rate_match_turbo with E = 3K + 12, rv 0 (rate_matching.py:163-229, core/ofdm_core.py:993-1001):
    out[i] = encoded[table[i]], -1 = constant 0 (padding of the shorter parity streams)."""
    enc = np.arange(3 * K + 12)
    d0 = np.concatenate([enc[0:3 * K:3], enc[3 * K:3 * K + 3], enc[3 * K + 6:3 * K + 9]])
    d1 = np.concatenate([enc[1:3 * K:3], enc[3 * K + 3:3 * K + 6]])
    d2 = np.concatenate([enc[2:3 * K:3], enc[3 * K + 9:3 * K + 12]])
    v = [d[sub_block_permutation(len(d))] for d in (d0, d1, d2)]
    n = max(len(x) for x in v)
    cb = np.full(3 * n, -1, dtype=np.int64)
    for j, x in enumerate(v):
        cb[j:3 * len(x):3] = x
    return cb[:3 * K + 12]


def rate_dematch_table(K):
    """rate_dematching_turbo (rate_matching.py:297-396): out[j] = llr[table[j]], -1 = 0.0 (never sent)."""
    src = rate_match_table(K)
    t = np.full(3 * K + 12, -1, dtype=np.int64)
    ok = src >= 0
    t[src[ok]] = np.flatnonzero(ok)
    return t


def llr_noise_var(H, snr_db, awgn_channel):
    """core/ofdm_core.py:1228-1250."""
    s2 = 1.0 / (10 ** (snr_db / 10))
    if awgn_channel:
        return np.full(len(H), s2)
    return np.maximum(s2 / np.clip(np.abs(H) ** 2, 1e-6, 1e6), s2 / 4.0)


def soft_demap(symbols, noise_var, modulation):
    """_calculate_llrs_qpsk/_16qam/_64qam (core/ofdm_core.py:791-923): QPSK exact and unclipped;
    16/64-QAM max-log over the natural-binary raster constellation, clipped to +-10."""
    y = np.asarray(symbols)
    nv = np.broadcast_to(np.asarray(noise_var, dtype=float), y.shape)
    if modulation == 'QPSK':
        out = np.zeros(2 * len(y))
        out[0::2] = (2.0 / nv) * y.real * np.sqrt(2)
        out[1::2] = (2.0 / nv) * y.imag * np.sqrt(2)
        return out
    c = constellation(modulation)
    b = BITS_PER_SYMBOL[modulation]
    bitmap = (np.arange(len(c))[:, None] >> (b - 1 - np.arange(b))[None, :]) & 1
    d = np.abs(y[:, None] - c[None, :]) ** 2
    out = np.zeros((len(y), b))
    for p in range(b):
        d0 = np.min(d[:, bitmap[:, p] == 0], axis=1)
        d1 = np.min(d[:, bitmap[:, p] == 1], axis=1)
        out[:, p] = np.clip((d1 - d0) / (2.0 * nv), -10.0, 10.0)
    return out.reshape(-1)


def _trellis():
    ns, so, po = np.zeros((8, 2), int), np.zeros((8, 2), int), np.zeros((8, 2), int)
    for st in range(8):
        s0, s1, s2 = (st >> 2) & 1, (st >> 1) & 1, st & 1
        for u in range(2):
            fb = (u + s1 + s2) % 2
            ns[st, u], so[st, u], po[st, u] = (fb << 2) | (s0 << 1) | s1, fb, (fb + s0 + s2) % 2
    return ns, so, po


def maxlog_bcjr(Ls, Lp, La, extrinsic=True, logmap=False):
    """LogMAPDecoder.decode (core/channel_coding/turbo_decoder.py:158-293), default max-log mode or
    (logmap=True) the exact max*(a, b) = log(e^a + e^b) of set_decoder_mode(False) (:64-120):
    start and end in state 0, gamma = (+-Ls +-Lp +-La)/2 with Ls signed by the FEEDBACK bit of the
    branch and La by its input bit."""
    if logmap:
        return _logmap_bcjr(Ls, Lp, La, extrinsic)
    ns, so, po = _trellis()
    n = len(Ls)
    g = ((1 - 2 * so)[None] * (Ls[:, None, None] / 2.0) + (1 - 2 * po)[None] * (Lp[:, None, None] / 2.0)) + \
        (1 - 2 * np.arange(2))[None, None, :] * (La[:, None, None] / 2.0)
    alpha = np.full((n + 1, 8), -np.inf)
    beta = np.full((n + 1, 8), -np.inf)
    alpha[0, 0] = beta[n, 0] = 0.0
    for k in range(n):
        cand = alpha[k][:, None] + g[k]
        nxt = np.full(8, -np.inf)
        np.maximum.at(nxt, ns.reshape(-1), cand.reshape(-1))
        alpha[k + 1] = nxt
    for k in range(n - 1, -1, -1):
        beta[k] = np.max(beta[k + 1][ns] + g[k], axis=1)
    val = (alpha[:n, :, None] + g) + beta[1:][:, ns]
    apost = np.max(val[:, :, 0], axis=1) - np.max(val[:, :, 1], axis=1)
    return (apost < 0).astype(np.uint8), (apost - La - Ls) if extrinsic else apost


def _logmap_bcjr(Ls, Lp, La, extrinsic):
    ns, so, po = _trellis()
    n = len(Ls)
    g = ((1 - 2 * so)[None] * (Ls[:, None, None] / 2.0) + (1 - 2 * po)[None] * (Lp[:, None, None] / 2.0)) + \
        (1 - 2 * np.arange(2))[None, None, :] * (La[:, None, None] / 2.0)
    alpha = np.full((n + 1, 8), -np.inf)
    beta = np.full((n + 1, 8), -np.inf)
    alpha[0, 0] = beta[n, 0] = 0.0
    with np.errstate(invalid='ignore'):
        for k in range(n):
            cand = alpha[k][:, None] + g[k]
            nxt = np.full(8, -np.inf)
            np.logaddexp.at(nxt, ns.reshape(-1), cand.reshape(-1))
            alpha[k + 1] = nxt
        for k in range(n - 1, -1, -1):
            beta[k] = np.logaddexp.reduce(beta[k + 1][ns] + g[k], axis=1)
        val = (alpha[:n, :, None] + g) + beta[1:][:, ns]
        apost = np.logaddexp.reduce(val[:, :, 0], axis=1) - np.logaddexp.reduce(val[:, :, 1], axis=1)
    return (apost < 0).astype(np.uint8), (apost - La - Ls) if extrinsic else apost


def turbo_decode(llr, K, f1, f2, num_iterations=8, logmap=False):
    """turbo_decode (turbo_decoder.py:340-446)."""
    pi = qpp_indices(K, f1, f2)
    inv = np.empty(K, dtype=np.int64)
    inv[pi] = np.arange(K)
    Ls = np.concatenate([llr[0:3 * K:3], llr[3 * K:3 * K + 3]])
    Lp1 = np.concatenate([llr[1:3 * K:3], llr[3 * K + 3:3 * K + 6]])
    Lp2 = np.concatenate([llr[2:3 * K:3], llr[3 * K + 9:3 * K + 12]])
    Ls2 = np.concatenate([Ls[:K][pi], llr[3 * K + 6:3 * K + 9]])
    e21 = np.zeros(K)
    z3 = np.zeros(3)
    for _ in range(num_iterations):
        _, e12 = maxlog_bcjr(Ls, Lp1, np.concatenate([e21, z3]), logmap=logmap)
        _, e21i = maxlog_bcjr(Ls2, Lp2, np.concatenate([e12[:K][pi], z3]), logmap=logmap)
        e21 = e21i[:K][inv]
    bits, _ = maxlog_bcjr(Ls, Lp1, np.concatenate([e21, z3]), extrinsic=False, logmap=logmap)
    return bits[:K]


def symbol_interleave(sym, ncols):
    """core/ofdm_core.py:1037-1060: write row-wise into [rows, ncols] (zero padded), read column-wise."""
    rows = int(np.ceil(len(sym) / ncols))
    m = np.zeros(rows * ncols, dtype=complex)
    m[:len(sym)] = sym
    return m.reshape(rows, ncols).T.reshape(-1), rows


def simulate_siso_coded(bits, snr_db, num, qpp, channel_type='awgn', itu_profile='Pedestrian_A',
                        frequency_ghz=2.0, velocity_kmh=0.0, phases=None, z=None, draws=None):
    """OFDMSimulator.simulate_siso_coded (core/ofdm_core.py:925-1338).  qpp: {K: (f1, f2)}."""
    bits = np.asarray(bits).astype(np.uint8)
    if bits.size == 0:
        raise ValueError("Bits array cannot be empty")
    data_idx, pilot_idx = grid_indices(num.N, num.Nc)
    Nd, b = len(data_idx), num.bits_per_symbol
    tb = np.concatenate([bits, crc24(bits, CRC24A_POLY)])
    blocks = segment_code_blocks(tb)
    coded = np.concatenate([turbo_encode(cb, *qpp[len(cb)])[np.maximum(rate_match_table(len(cb)), 0)] *
                            (rate_match_table(len(cb)) >= 0) for cb in blocks]).astype(np.uint8)
    sym = qam_map(coded, num.modulation)
    inter, rows = symbol_interleave(sym, Nd)
    sig = ofdm_modulate_grid(map_grid(inter.reshape(rows, Nd), num), num).reshape(-1)
    papr_db, papr_lin = papr(sig)
    if draws is None and (phases is None or z is None):
        draws = ReferenceDraws(len(pilot_idx))
    rx, _ = channel_link(sig, num, channel_type, snr_db, draws, itu_profile, frequency_ghz, velocity_kmh,
                         phases=phases, z=z)
    Y = rx_fft_stream(rx, num)
    H = estimate_channel_periodic(Y, num)
    eq = zf_equalize(Y, H)[:, data_idx].reshape(-1)
    Hd = H[:, data_idx].reshape(-1)
    nsym = len(coded) // b
    rrows = int(np.ceil(nsym / Nd))
    tot = rrows * Nd
    eq = np.pad(eq, (0, max(0, tot - len(eq))))[:tot]
    Hd = np.pad(Hd, (0, max(0, tot - len(Hd))), mode='edge')[:tot]
    sym_rx = eq.reshape(Nd, rrows).T.reshape(-1)[:nsym]
    H_rx = Hd.reshape(Nd, rrows).T.reshape(-1)[:nsym]
    nv = llr_noise_var(H_rx, snr_db, channel_type != 'rayleigh_mp')
    llr = soft_demap(sym_rx, nv, num.modulation)[:len(coded)]
    dec, off = [], 0
    for cb in blocks:
        K = len(cb)
        blk = llr[off:off + 3 * K + 12]
        off += 3 * K + 12
        t = rate_dematch_table(K)
        dec.append(turbo_decode(np.where(t >= 0, blk[np.maximum(t, 0)], 0.0), K, *qpp[K]))
    tb_rx = desegment_code_blocks(dec, len(tb))
    crc_ok = bool(np.array_equal(tb_rx[-24:], crc24(tb_rx[:-24], CRC24A_POLY)))
    bits_rx = tb_rx[:-24][:len(bits)]
    errors = int(np.sum(bits != bits_rx))
    return dict(signal_tx=sig, signal_rx=rx, coded_bits=coded, symbols_tx=sym, symbols_rx=sym_rx, H_estimate=H_rx,
                llr=llr, noise_var_mean=float(np.mean(nv)), bits_rx=bits_rx, errors=errors, ber=errors / len(bits),
                crc_pass=crc_ok, coded_bits_length=len(coded), papr_db=papr_db, papr_linear=papr_lin,
                decoded_blocks=dec, code_blocks=blocks)
