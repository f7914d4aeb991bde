"""Host-side table work of the link chain (runs once per configuration).

Everything here is integer / table arithmetic that the reference does in Python
before the sample-level work starts; the results are uploaded into an `lte_plan`.
"""
import numpy as np

from config import ITU_CHANNEL_MODELS

from . import _native as nat


def grid_indices(N, Nc):
    """Data / pilot bins in raw FFT-bin order (reference core/resource_mapper.py:45-74)."""
    gl = (N - Nc) // 2
    gr = N - Nc - gl
    k = np.arange(N)
    useful = (k >= gl) & (k < N - gr) & (k != N // 2)
    pilot = useful & ((k - gl) % 6 == 3)
    return np.flatnonzero(useful & ~pilot), np.flatnonzero(pilot)


def pilot_values(cell_id, num_pilots):
    """(+-1)(1+1j)/sqrt(2) with signs from NumPy's legacy MT19937 stream seeded by the
    cell id (reference core/resource_mapper.py:137-152).  Uses a private RandomState,
    so the caller's global RNG is left alone."""
    signs = np.random.RandomState(cell_id).choice([1, -1], size=num_pilots)
    return signs * ((1 + 1j) / np.sqrt(2))


def mimo_pilot_sets(num_tx, num_pilots):
    """Frequency-interleaved pilot ownership for N TX antennas (reference
    core/mimo_channel_estimator_periodic.py:75-106): TX t owns pilot_idx[t % step::step],
    step = min(num_tx, 4), sign stream of cell t % 4; zero elsewhere."""
    step = min(num_tx, 4)
    sets = np.zeros((num_tx, num_pilots), dtype=complex)
    for t in range(num_tx):
        own = np.arange(t % step, num_pilots, step)
        vals = pilot_values(t % 4, len(own))
        sets[t, own] = vals
    return sets


def doppler_hz(frequency_ghz, velocity_kmh):
    """reference core/channel.py:118-143."""
    return ((velocity_kmh / 3.6) * (frequency_ghz * 1e9)) / 3e8


def channel_desc(channel_type, fs, itu_profile='Pedestrian_A', frequency_ghz=2.0, velocity_kmh=0.0,
                 gain_conversions=2, faithful_gains=True):
    """Tapped-delay-line description.  `faithful_gains` keeps the reference's repeated
    dB->linear conversion (core/channel.py:184 then core/rayleighchannel.py:16, and a
    third time on the spatial-multiplexing path, core/channel.py:436)."""
    d = nat.ChannelDesc()
    if channel_type != 'rayleigh_mp':
        d.num_taps = 0
        return d
    if itu_profile not in ITU_CHANNEL_MODELS:
        raise ValueError(f"Perfil ITU no encontrado: {itu_profile}. "
                         f"Opciones disponibles: {list(ITU_CHANNEL_MODELS.keys())}")
    prof = ITU_CHANNEL_MODELS[itu_profile]
    g = np.array(prof['power_db'], dtype=float)
    for _ in range(gain_conversions if faithful_gains else 1):
        g = 10 ** (g / 20)
    delays = np.array(prof['delays_us']) * 1e-6
    d.num_taps = len(g)
    for i in range(len(g)):
        d.delay[i] = int(np.round(delays[i] * fs))
        d.gain[i] = float(g[i])
    if frequency_ghz is None or velocity_kmh is None:     # reference core/channel.py:121-139
        if 'Pedestrian' in itu_profile:
            velocity_kmh = 5.0
        elif 'Vehicular_A' in itu_profile:
            velocity_kmh = 30.0
        elif 'Vehicular_B' in itu_profile:
            velocity_kmh = 120.0
        else:
            velocity_kmh = 10.0
        frequency_ghz = 2.0
    d.doppler_hz = float(doppler_hz(frequency_ghz, velocity_kmh))
    return d
