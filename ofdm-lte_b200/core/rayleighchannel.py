"""Sum-of-16-sinusoids Jakes fading and the time-domain tapped delay line
(reference core/rayleighchannel.py:5-109), evaluated by csrc/channel.cu."""
import ctypes as C

import numpy as np
import torch

from lte_b200 import _native as nat
from lte_b200 import tables

from . import _backend as be


class _FsConfig:
    """Plan carrier for a stand-alone channel object (only fs matters to the TDL kernel)."""
    N, Nc, cp_length, bits_per_symbol = 128, 76, 9, 2

    def __init__(self, fs):
        self.fs = fs


class RayleighChannel:
    def __init__(self, Fs, fD, delays, gains):
        """Fs [Hz], fD [Hz], delays [s], gains: the reference converts them with 10**(g/20)
        whatever their unit (core/rayleighchannel.py:16); kept."""
        self.Fs = Fs
        self.fD = fD
        self.delays = np.array(delays)
        self.gains = 10 ** (np.array(gains) / 20)
        assert len(self.delays) == len(self.gains), "delays y gains deben tener la misma longitud"
        self.num_paths = len(delays)

    def _desc(self):
        d = nat.ChannelDesc()
        d.num_taps = self.num_paths
        for i in range(self.num_paths):
            d.delay[i] = int(np.round(self.delays[i] * self.Fs))
            d.gain[i] = float(self.gains[i])
        d.doppler_hz = float(self.fD)
        return d

    def _filter_device(self, x_t, phases_u):
        """x_t complex64 CUDA [1, n]; phases_u float32 CUDA [taps, 16] in [0, 1)."""
        eng = be.engine_for(_FsConfig(self.Fs))
        faded, power = eng.channel(x_t, self._desc(), 1, 1, phases=phases_u.reshape(1, -1))
        return faded.reshape(1, -1), power

    def filter(self, x):
        x_t = be.as_complex_tensor(x).reshape(1, -1)
        faded, _ = self._filter_device(x_t, be.NumpyDraws().phases(self.num_paths))
        return be.to_numpy(faded.reshape(-1))

    def jakes_fading(self, N, N_s=16):
        """h[m] for m = 0..N-1 of one tap (unit gain), by filtering a constant 1."""
        if N_s != 16:
            raise ValueError("the CUDA Jakes generator uses 16 sinusoids (reference default N_s=16)")
        one = RayleighChannel(self.Fs, self.fD, [0.0], [0.0])
        x_t = torch.ones((1, N), dtype=torch.complex64, device=be.device())
        faded, _ = one._filter_device(x_t, be.NumpyDraws().phases(1))
        return be.to_numpy(faded.reshape(-1))

    def impulse_response(self, N=1):
        taps = [self.gains[i] * self.jakes_fading(max(N, 1))[0] for i in range(self.num_paths)]
        return np.array(self.delays), np.array(taps)

    def channel_response(self, freqs, h_taps, N_freq=None):
        Hf = np.zeros_like(freqs, dtype=complex)
        for i in range(self.num_paths):
            Hf += h_taps[i] * np.exp(-1j * 2 * np.pi * freqs * self.delays[i])
        return Hf
