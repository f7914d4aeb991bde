"""Beamforming precoders -- the reference's core/beamforming_precoder.py:16-292.

MRT weights come from the `lte_bf_weights` kernel (the B = 1 case of the batched sweep path); applying a
rank-1 precoder is the L = 1 case of the layer-precoding kernel `lte_sm_precode`."""
import numpy as np
import torch

from . import _backend as be
from .csi_feedback import _as_batch, _bf_engine

_ROW = 1024                 # data positions per plan row when a symbol vector is longer than one plan (<= 2048 bins)


class _Bins:
    """simple-mode plan with exactly n data positions."""
    cp_length, fs, bits_per_symbol = 0, 1.92e6, 2

    def __init__(self, n):
        self.Nc = int(n)
        self.N = max(64, 1 << (int(n) - 1).bit_length())


def _unit(vec):
    return vec / np.sqrt(np.vdot(vec, vec).real)


class BeamformingPrecoder:
    def __init__(self, num_tx, num_layers=1, precoder_type='MRT'):
        self.num_tx, self.num_layers, self.precoder_type = num_tx, num_layers, precoder_type
        self.W = None

    # ---- weight computation -------------------------------------------------------------------------
    def calculate_mrt_weights(self, H_channel):
        """conj(mean over RX of H) normalised to unit power -> [num_tx, 1]."""
        W = _bf_engine().bf_weights(_as_batch(H_channel), None, mode='MRT')[0]
        return be.to_numpy(W[0]).astype(complex)[:, None]

    def calculate_eigenbeamforming(self, H_channel):
        """Dominant eigenvector of H^H H (T x T host eigen-decomposition; not on the simulate_beamforming path)."""
        H = np.asarray(H_channel)
        lam, vec = np.linalg.eig(H.conj().T @ H)
        return _unit(vec[:, np.argmax(np.abs(lam))])[:, None]

    def update_precoder(self, H_channel, method='MRT'):
        H = np.asarray(H_channel)
        if H.ndim == 3:                                # [R, T, subcarriers]: average over frequency first
            H = H.mean(axis=2)
        rule = {'MRT': self.calculate_mrt_weights, 'eigen': self.calculate_eigenbeamforming}.get(method)
        if rule is None:
            raise ValueError(f"Método '{method}' no soportado")
        self.W = rule(H)
        return self.W

    def get_current_precoder(self):
        return self.W

    # ---- use of the weights --------------------------------------------------------------------------
    def _need_W(self, msg):
        if self.W is None:
            raise ValueError(msg)
        return self.W

    def apply_precoding(self, symbols, W_matrix=None):
        """x = W s: symbols [n] or [layers, n] -> [num_tx, n]."""
        W = np.asarray(self._need_W("Precoder W no ha sido calculado. Llamar a update_precoder() primero.")
                       if W_matrix is None else W_matrix)
        s = np.atleast_2d(np.asarray(symbols.cpu() if isinstance(symbols, torch.Tensor) else symbols))
        if W.shape[1] != 1 or s.shape[0] != 1:
            return W @ s                               # multi-layer precoding is LayerMapper + lte_sm_precode (SM path)
        n = s.shape[1]
        per_row = n if n <= 2 * _ROW else _ROW
        rows = -(-n // per_row)
        padded = np.zeros(rows * per_row, dtype=np.complex64)
        padded[:n] = s[0]
        eng = be.engine_for(_Bins(per_row), mode='simple')
        out, _ = eng.sm_precode(rows, W, symbols=be.as_complex_tensor(padded[None]))
        return be.to_numpy(out)[:, :n].astype(complex)

    def get_effective_channel(self, H_channel):
        return np.asarray(H_channel) @ self._need_W("Precoder W no disponible")

    def calculate_beamforming_gain(self, H_channel):
        """10 log10(||H W||^2 / (||H||_F^2 / T)); 0.0 while no precoder has been set."""
        if self.W is None:
            return 0.0
        H = np.asarray(H_channel)
        return float(10 * np.log10(np.linalg.norm(H @ self.W) ** 2 / (np.linalg.norm(H) ** 2 / self.num_tx)))


class AdaptiveBeamforming(BeamformingPrecoder):
    """Precoder refreshed once per tenth of the coherence time (reference :204-292)."""

    SYMBOL_S = 1.0 / 15000          # OFDM symbol duration at 15 kHz spacing

    def __init__(self, num_tx, velocity_kmh, frequency_ghz, num_layers=1):
        super().__init__(num_tx, num_layers, precoder_type='MRT')
        self.velocity_kmh, self.frequency_ghz = velocity_kmh, frequency_ghz
        self.update_period = self._calculate_update_period()
        self.symbols_since_update = 0

    def _calculate_update_period(self):
        doppler = self.velocity_kmh / 3.6 * self.frequency_ghz * 1e9 / 3e8
        if doppler == 0:
            return 100
        coherence = 9 / (16 * np.pi * doppler)
        return np.clip(int(0.1 * coherence / self.SYMBOL_S), 1, 140)

    def should_update(self):
        return self.symbols_since_update >= self.update_period

    def process_symbol(self, symbols, H_channel):
        if self.W is None or self.should_update():
            self.update_precoder(H_channel, method='MRT')
            self.symbols_since_update = 0
        self.symbols_since_update += 1
        return self.apply_precoding(symbols)
