"""Beamforming precoders -- reference core/beamforming_precoder.py:16-292.  MRT weights, the
effective channel and the beamforming gain come from `lte_bf_weights`; `apply_precoding` is the
rank-1 case of the layer-precoding kernel (`lte_sm_precode`)."""
import numpy as np
import torch

from . import _backend as be
from .csi_feedback import _bf_engine


class BeamformingPrecoder:
    def __init__(self, num_tx, num_layers=1, precoder_type='MRT'):
        self.num_tx = num_tx
        self.num_layers = num_layers
        self.precoder_type = precoder_type
        self.W = None

    @staticmethod
    def _h_device(H_channel):
        H = np.asarray(H_channel, dtype=np.complex64)
        if H.ndim == 1:
            H = H.reshape(1, -1)
        return be.as_complex_tensor(H[None])

    def calculate_mrt_weights(self, H_channel):
        """W = conj(mean_r H) / ||.||  (reference :41-66) -> [num_tx, 1]."""
        W, _, _, _ = _bf_engine().bf_weights(self._h_device(H_channel), None, mode='MRT')
        return be.to_numpy(W[0]).astype(complex).reshape(-1, 1)

    def calculate_eigenbeamforming(self, H_channel):
        """Dominant eigenvector of H^H H (reference :68-93); a T x T host eigen-decomposition, not on
        the simulate_beamforming path."""
        H = np.asarray(H_channel)
        ev, vec = np.linalg.eig(H.conj().T @ H)
        W = vec[:, np.argmax(np.abs(ev))]
        return (W / np.sqrt(np.sum(np.abs(W) ** 2))).reshape(-1, 1)

    def apply_precoding(self, symbols, W_matrix=None):
        """x = W s (reference :95-127): symbols [n] or [layers, n] -> [num_tx, n]."""
        if W_matrix is None:
            if self.W is None:
                raise ValueError("Precoder W no ha sido calculado. Llamar a update_precoder() primero.")
            W_matrix = self.W
        W_matrix = np.asarray(W_matrix)
        sym = np.asarray(symbols.cpu() if isinstance(symbols, torch.Tensor) else symbols)
        if sym.ndim == 1:
            sym = sym.reshape(1, -1)
        if W_matrix.shape[1] != 1 or sym.shape[0] != 1:
            # multi-layer precoding goes through LayerMapper + lte_sm_precode in the SM path
            return W_matrix @ sym
        n = sym.shape[1]
        nd = n if n <= 2048 else 1024                     # plan sizes stop at 2048 bins: longer inputs go in rows
        S = -(-n // nd)
        buf = np.zeros(S * nd, dtype=np.complex64)
        buf[:n] = sym.reshape(-1)
        eng = be.engine_for(_Bins(nd), mode='simple')
        out, _ = eng.sm_precode(S, W_matrix, symbols=be.as_complex_tensor(buf.reshape(1, -1)))
        return be.to_numpy(out)[:, :n].astype(complex)

    def update_precoder(self, H_channel, method='MRT'):
        H = np.asarray(H_channel)
        H_avg = np.mean(H, axis=2) if H.ndim == 3 else H
        if method == 'MRT':
            self.W = self.calculate_mrt_weights(H_avg)
        elif method == 'eigen':
            self.W = self.calculate_eigenbeamforming(H_avg)
        else:
            raise ValueError(f"Método '{method}' no soportado")
        return self.W

    def get_current_precoder(self):
        return self.W

    def get_effective_channel(self, H_channel):
        if self.W is None:
            raise ValueError("Precoder W no disponible")
        return np.asarray(H_channel) @ self.W

    def calculate_beamforming_gain(self, H_channel):
        """10 log10(||H W||^2 / (||H||_F^2 / T)); 0.0 while no precoder is set (reference :176-201)."""
        if self.W is None:
            return 0.0
        H = np.asarray(H_channel)
        return float(10 * np.log10(np.sum(np.abs(H @ self.W) ** 2) / (np.sum(np.abs(H) ** 2) / self.num_tx)))


class _Bins:
    """simple-mode plan with exactly n data positions."""
    cp_length, fs, bits_per_symbol = 0, 1.92e6, 2

    def __init__(self, n):
        self.Nc = int(n)
        self.N = max(64, 1 << (int(n) - 1).bit_length())


class AdaptiveBeamforming(BeamformingPrecoder):
    """Periodic precoder refresh from the coherence time (reference :204-292)."""

    def __init__(self, num_tx, velocity_kmh, frequency_ghz, num_layers=1):
        super().__init__(num_tx, num_layers, precoder_type='MRT')
        self.velocity_kmh = velocity_kmh
        self.frequency_ghz = frequency_ghz
        self.update_period = self._calculate_update_period()
        self.symbols_since_update = 0

    def _calculate_update_period(self):
        fd = (self.velocity_kmh / 3.6) * (self.frequency_ghz * 1e9) / 3e8
        if fd == 0:
            return 100
        Tc = 9 / (16 * np.pi * fd)
        return np.clip(int(0.1 * Tc / (1 / 15000)), 1, 140)

    def should_update(self):
        return self.symbols_since_update >= self.update_period

    def process_symbol(self, symbols, H_channel):
        if self.should_update() or self.W is None:
            self.update_precoder(H_channel, method='MRT')
            self.symbols_since_update = 0
        tx = self.apply_precoding(symbols)
        self.symbols_since_update += 1
        return tx
