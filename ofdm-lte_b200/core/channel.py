"""AWGN and ITU-R M.1225 Rayleigh multipath channels (reference core/channel.py).

AWGN is added by `lte_awgn_add` with sigma taken from the measured stream power exactly as the
reference does (core/channel.py:46-66, :216-232).  Draws come from NumPy's legacy global RNG in
the reference's order by default, or from Philox when the owner asks for rng='philox'."""
import numpy as np
import torch

from config import ITU_CHANNEL_MODELS
from lte_b200 import tables

from . import _backend as be
from .rayleighchannel import RayleighChannel, _FsConfig


def _awgn_device(x_t, snr_linear, draws, fs_cfg, row_id0=0, power=None):
    """x_t complex64 CUDA [rows, n]: y = x + sigma z, sigma from each row's measured power
    (sum |x|^2, produced by the channel kernel or by the power kernel of lte_channel_tdl)."""
    eng = be.engine_for(fs_cfg)
    rows, n = x_t.shape
    if power is None:
        _, power = eng.channel(x_t, tables.channel_desc('awgn', fs_cfg.fs), rows, 1)
    power = power.reshape(-1)
    snr = torch.full((rows,), float(snr_linear), dtype=torch.float32, device=x_t.device)
    if draws.kind == 'numpy':
        z = torch.stack([draws.unit_normals(n) for _ in range(rows)])
        y = eng.awgn(x_t, 1, power, snr, rows, z=z)
    else:
        y = eng.awgn(x_t, 1, power, snr, rows, seed=draws.seed, row_id0=row_id0)
    noise_power = float(power[0].item() / n / snr_linear)
    return y, noise_power


class AWGNChannel:
    def __init__(self, snr_db=10.0):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)
        self.noise_power = None
        self._draws = be.NumpyDraws()

    def set_snr(self, snr_db):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)

    def _transmit_device(self, x_t, row_id0=0):
        y, self.noise_power = _awgn_device(x_t, self.snr_linear, self._draws, _FsConfig(1.92e6), row_id0)
        return y

    def transmit(self, signal):
        x_t = be.as_complex_tensor(signal).reshape(1, -1)
        y = self._transmit_device(x_t)
        out = be.to_numpy(y.reshape(-1))
        return out, out - np.asarray(signal, dtype=np.complex64)

    def get_noise_power(self):
        return self.noise_power

    def get_snr_info(self):
        return {'SNR (dB)': self.snr_db, 'SNR (lineal)': self.snr_linear, 'Potencia de ruido': self.noise_power}


class RayleighMultiPathChannel:
    """ITU profile wrapper: computes fD and the (double-converted) tap gains
    (reference core/channel.py:83-245)."""

    def __init__(self, snr_db=10.0, fs=None, itu_profile='Vehicular_A', fD=None,
                 frequency_ghz=None, velocity_kmh=None, verbose=True):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)
        self.itu_profile = itu_profile
        self.fs = fs
        self.noise_power = None
        self.frequency_ghz = frequency_ghz
        self.velocity_kmh = velocity_kmh
        self.verbose = verbose
        delays, gains = self._get_itu_profile_params(itu_profile)
        if fD is None:
            fD = tables.channel_desc('rayleigh_mp', fs, itu_profile, frequency_ghz, velocity_kmh).doppler_hz
        self.rayleigh = RayleighChannel(fs, fD, delays, gains)
        self._draws = be.NumpyDraws()
        if self.verbose:
            print(f"[RayleighMultiPathChannel] Perfil: {itu_profile}")
            print(f"  - Doppler máximo: {fD:.1f} Hz")

    def _get_itu_profile_params(self, profile_name):
        if profile_name not in ITU_CHANNEL_MODELS:
            raise ValueError(f"Perfil ITU no encontrado: {profile_name}. "
                             f"Opciones disponibles: {list(ITU_CHANNEL_MODELS.keys())}")
        prof = ITU_CHANNEL_MODELS[profile_name]
        return np.array(prof['delays_us']) * 1e-6, 10 ** (np.array(prof['power_db']) / 20)

    def set_snr(self, snr_db):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)

    def set_profile(self, itu_profile):
        self.itu_profile = itu_profile
        delays, gains = self._get_itu_profile_params(itu_profile)
        self.rayleigh.delays = np.array(delays)
        self.rayleigh.gains = gains
        self.rayleigh.num_paths = len(delays)

    def _transmit_device(self, x_t, row_id0=0):
        if self._draws.kind == 'numpy':
            u = self._draws.phases(self.rayleigh.num_paths)
        else:
            eng = be.engine_for(_FsConfig(self.fs))
            u = eng.random_phases(1, self.rayleigh.num_paths * 16, self._draws.seed, row_id0)
        faded, power = self.rayleigh._filter_device(x_t, u)
        y, self.noise_power = _awgn_device(faded, self.snr_linear, self._draws, _FsConfig(self.fs), row_id0,
                                           power=power)
        return y

    def transmit(self, signal):
        x_t = be.as_complex_tensor(signal).reshape(1, -1)
        return be.to_numpy(self._transmit_device(x_t).reshape(-1)), None

    def get_channel_info(self):
        return {'type': 'Rayleigh MultiPath (ITU-R M.1225)', 'profile': self.itu_profile, 'SNR_dB': self.snr_db,
                'num_paths': self.rayleigh.num_paths, 'delays_us': self.rayleigh.delays * 1e6,
                'gains_dB': 20 * np.log10(self.rayleigh.gains)}


class FadingChannel:
    """Flat per-sample Rayleigh fading + AWGN (reference core/channel.py:248-291)."""

    def __init__(self, snr_db=10.0, fading_type='rayleigh'):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)
        self.fading_type = fading_type
        self.awgn_channel = AWGNChannel(snr_db)

    def set_snr(self, snr_db):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)
        self.awgn_channel.set_snr(snr_db)

    def transmit(self, signal):
        n = len(signal)
        h = np.random.normal(0, 1 / np.sqrt(2), n) + 1j * np.random.normal(0, 1 / np.sqrt(2), n)
        x_t = be.as_complex_tensor(signal).reshape(1, -1) * be.as_complex_tensor(h).reshape(1, -1)
        y = self.awgn_channel._transmit_device(x_t)
        return be.to_numpy(y.reshape(-1)), h


class ChannelSimulator:
    """Dispatcher over the channel types (reference core/channel.py:294-395)."""

    def __init__(self, channel_type='awgn', snr_db=10.0, fs=None, itu_profile='Vehicular_A',
                 frequency_ghz=None, velocity_kmh=None, verbose=True):
        self.channel_type = channel_type
        self.fs = fs
        self.itu_profile = itu_profile
        self.frequency_ghz = frequency_ghz
        self.velocity_kmh = velocity_kmh
        if channel_type == 'awgn':
            self.channel = AWGNChannel(snr_db)
        elif channel_type == 'fading':
            self.channel = FadingChannel(snr_db)
        elif channel_type == 'rayleigh_mp':
            if fs is None:
                raise ValueError("Se requiere fs (frecuencia de muestreo) para canal Rayleigh")
            self.channel = RayleighMultiPathChannel(snr_db, fs, itu_profile, frequency_ghz=frequency_ghz,
                                                    velocity_kmh=velocity_kmh, verbose=verbose)
        else:
            raise ValueError(f"Tipo de canal desconocido: {channel_type}")

    def _set_draws(self, draws):
        ch = self.channel.awgn_channel if isinstance(self.channel, FadingChannel) else self.channel
        ch._draws = draws

    def _transmit_device(self, x_t, row_id0=0):
        if isinstance(self.channel, FadingChannel):
            y, _ = self.channel.transmit(be.to_numpy(x_t.reshape(-1)))
            return be.as_complex_tensor(y).reshape(1, -1)
        return self.channel._transmit_device(x_t, row_id0)

    def transmit(self, signal):
        received, _ = self.channel.transmit(signal)
        return received

    def set_snr(self, snr_db):
        self.channel.set_snr(snr_db)

    # -- spatial multiplexing channel (reference core/channel.py:397-493) ----------------------
    def _transmit_sm_device(self, eng, tx_t, num_rx, draws):
        """tx_t [T, n] -> (rx [R, n], H_channel [R, T] numpy).  Multipath: R*T independent
        RayleighChannel links (tap gains converted a third time, :436,444) summed per RX; flat:
        one CN(0,1) coefficient per link.  One AWGN per RX from that antenna's measured power."""
        T, n = tx_t.shape
        dev = tx_t.device
        snr_lin = float(self.channel.snr_linear)
        snr = torch.full((num_rx,), snr_lin, dtype=torch.float32, device=dev)
        if self.channel_type == 'rayleigh_mp':
            ray = self.channel.rayleigh
            chan = tables.channel_desc('rayleigh_mp', self.fs, self.itu_profile, self.frequency_ghz,
                                       self.velocity_kmh, gain_conversions=3)
            chan.doppler_hz = float(ray.fD)
            taps = chan.num_taps
            Hc = np.zeros((num_rx, T), dtype=complex)
            if draws.kind == 'numpy':
                # per (rx, tx): taps x rand(16) for filter(), then taps x rand(16) for impulse_response(N=1)
                u = np.zeros((num_rx, T, taps, 16))
                g0 = float(10 ** (ray.gains[0] / 20))
                for r in range(num_rx):
                    for t in range(T):
                        u[r, t] = np.stack([np.random.rand(16) for _ in range(taps)])
                        u2 = np.stack([np.random.rand(16) for _ in range(taps)])
                        Hc[r, t] = g0 * np.sqrt(2 / 16) * np.sum(np.exp(2j * np.pi * u2[0]))   # jakes_fading(1)[0]
                ut = torch.from_numpy(u.astype(np.float32)).to(dev).reshape(1, -1)
            else:
                sid = draws.next_stream()
                ut = eng.random_phases(1, num_rx * T * taps * 16, draws.seed, sid)
                # the reported channel matrix is the first tap's Jakes sum at sample 0 (reference impulse_response(N=1)):
                # with Philox draws it comes from the link's own phases instead of a second, unrelated draw
                u0 = be.to_numpy(ut.reshape(num_rx, T, taps, 16)[:, :, 0, :]).astype(np.float64)
                Hc = float(10 ** (ray.gains[0] / 20)) * np.sqrt(2 / 16) * np.sum(np.exp(2j * np.pi * u0), axis=-1)
            faded, power = eng.channel(tx_t, chan, 1, num_rx, T=T, phases=ut)
            acc = faded.reshape(num_rx, n)
        else:
            if draws.kind == 'numpy':
                h = np.zeros((num_rx, T), dtype=complex)
                for r in range(num_rx):
                    for t in range(T):
                        h[r, t] = np.random.normal(0, 1 / np.sqrt(2)) + 1j * np.random.normal(0, 1 / np.sqrt(2))
            else:
                g = torch.Generator(device='cpu').manual_seed(draws.seed * 1000003 + draws.next_stream())
                hh = torch.randn((num_rx, T, 2), generator=g, dtype=torch.float64) / np.sqrt(2)
                h = (hh[..., 0] + 1j * hh[..., 1]).numpy()
            Hc = h
            acc, power = eng.flat_mimo(tx_t, be.as_complex_tensor(h).reshape(1, num_rx, T), 1, num_rx, T)
        if draws.kind == 'numpy':
            z = torch.stack([draws.unit_normals(n) for _ in range(num_rx)])
            rx = eng.awgn(acc, 1, power.reshape(-1), snr, num_rx, z=z)
        else:
            sid2 = draws.next_stream()
            rx = eng.awgn(acc, 1, power.reshape(-1), snr, num_rx, seed=draws.seed, row_id0=sid2 * num_rx)
        return rx, Hc

    def transmit_spatial_multiplexing(self, tx_signals, num_rx=2):
        n = min(len(sig) for sig in tx_signals)
        tx_t = torch.stack([be.as_complex_tensor(sig).reshape(-1)[:n] for sig in tx_signals])
        eng = be.engine_for(_FsConfig(self.fs if self.fs is not None else 1.92e6))
        rx, Hc = self._transmit_sm_device(eng, tx_t, num_rx, be.NumpyDraws())
        out = be.to_numpy(rx)
        return [out[r] for r in range(num_rx)], Hc

    def set_channel_type(self, channel_type, **kwargs):
        snr_db = getattr(self.channel, 'snr_db', 10.0)
        self.channel_type = channel_type
        if channel_type == 'awgn':
            self.channel = AWGNChannel(snr_db)
        elif channel_type == 'fading':
            self.channel = FadingChannel(snr_db)
        elif channel_type == 'rayleigh_mp':
            if self.fs is None:
                raise ValueError("Se requiere fs para cambiar a canal Rayleigh")
            self.itu_profile = kwargs.get('itu_profile', self.itu_profile)
            self.channel = RayleighMultiPathChannel(snr_db, self.fs, self.itu_profile)
        else:
            raise ValueError(f"Tipo de canal desconocido: {channel_type}")

    def set_itu_profile(self, itu_profile):
        if isinstance(self.channel, RayleighMultiPathChannel):
            self.channel.set_profile(itu_profile)
            self.itu_profile = itu_profile
        else:
            raise ValueError("Solo se puede cambiar perfil ITU en canal Rayleigh")

    def get_channel(self):
        return self.channel

    def get_channel_info(self):
        if hasattr(self.channel, 'get_channel_info'):
            return self.channel.get_channel_info()
        return {'type': self.channel_type, 'SNR_dB': getattr(self.channel, 'snr_db', None)}
