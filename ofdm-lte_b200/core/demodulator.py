"""OFDM demodulator façade and hard symbol detector (reference core/demodulator.py)."""
import numpy as np
import torch

from lte_b200 import _native as nat

from . import _backend as be
from .dft_precoding import SC_FDMDecodifier
from .modulator import QAMModulator


class OFDMDemodulator:
    """mode 'lte': LTEReceiver chain; mode 'simple': FFT and the first Nc bins
    (reference core/demodulator.py:15-188)."""

    def __init__(self, config, mode='simple', enable_equalization=False, enable_sc_fdm=False):
        self.config = config
        self.mode = mode
        self.enable_equalization = enable_equalization
        self.enable_sc_fdm = enable_sc_fdm
        self.qam_demodulator = QAMModulator(config.modulation)
        self.resource_mapper = None
        self.sc_fdm_decoder = None
        if mode == 'lte' or enable_sc_fdm:
            from .resource_mapper import ResourceMapper
            self.resource_mapper = ResourceMapper(config)
            if enable_sc_fdm:
                self.sc_fdm_decoder = SC_FDMDecodifier(len(self.resource_mapper.get_data_indices()), enable=True)
        if mode == 'lte':
            from .lte_receiver import LTEReceiver
            self.lte_receiver = LTEReceiver(config, cell_id=0, enable_equalization=enable_equalization,
                                            enable_sc_fdm=enable_sc_fdm)
        else:
            self.lte_receiver = None

    # -- device-level core ---------------------------------------------------------------------
    def _demodulate_stream_device(self, rx_t):
        """rx_t complex64 CUDA [1, n] -> data symbols [1, nsym] (before the slicer)."""
        if self.mode == 'lte' and self.lte_receiver is not None:
            return self.lte_receiver._receive_device(rx_t)['data']
        eng = be.engine_for(self.config, mode='simple')
        n = rx_t.shape[1]
        S = int(np.ceil(n / eng.L))
        if n < S * eng.L:
            rx_t = torch.nn.functional.pad(rx_t, (0, S * eng.L - n))
        Y = eng.rx_fft(rx_t, 1, S, nat.WINDOW_USEFUL)          # bins 0..Nc-1 (reference :116)
        return Y.reshape(1, -1)

    def demodulate(self, received_signal):
        """One OFDM symbol (reference :68-118)."""
        rx = be.as_complex_tensor(received_signal).reshape(1, -1)
        L = self.config.N + self.config.cp_length
        rx = torch.nn.functional.pad(rx, (0, L - rx.shape[1])) if rx.shape[1] < L else rx[:, :L].contiguous()
        if self.enable_sc_fdm and self.sc_fdm_decoder is not None and self.resource_mapper is not None:
            eng = be.engine_for(self.config)
            data = eng.zf(eng.rx_fft(rx, 1, 1, nat.WINDOW_FULL), None, 1, 1, nat.WINDOW_FULL)
            out = eng.dft_m(data, eng.Nd, inverse=True).reshape(-1)
        else:
            eng = be.engine_for(self.config, mode='simple')
            out = eng.rx_fft(rx, 1, 1, nat.WINDOW_USEFUL).reshape(-1)
        return be.to_numpy(out)

    def demodulate_stream(self, received_signal, num_ofdm_symbols=None, resource_grid=None):
        """-> (data symbols, bits) (reference :120-184)."""
        rx = be.as_complex_tensor(received_signal).reshape(1, -1)
        if num_ofdm_symbols is not None and self.mode != 'lte':
            L = self.config.N + self.config.cp_length
            need = int(num_ofdm_symbols) * L
            rx = torch.nn.functional.pad(rx, (0, need - rx.shape[1])) if rx.shape[1] < need else rx[:, :need].contiguous()
        data = self._demodulate_stream_device(rx)
        eng = be.engine_for(self.config, mode='lte' if self.mode == 'lte' else 'simple')
        _, idx = eng.demap_count(data, want_idx=True)
        bits = eng.indices_to_bits(idx, idx.shape[1] * eng.bps).reshape(-1)
        return be.to_numpy(data.reshape(-1)), be.to_numpy(bits, np.int64)

    def get_qam_demodulator(self):
        return self.qam_demodulator


class SymbolDetector:
    """Nearest constellation point, first minimum wins (reference core/demodulator.py:191-245)."""

    def __init__(self, constellation):
        self.constellation = np.asarray(constellation)
        n = len(self.constellation)
        self._mod = {4: 'QPSK', 16: '16-QAM', 64: '64-QAM'}.get(n)
        if self._mod is None or not np.allclose(self.constellation, QAMModulator(self._mod).constellation):
            raise ValueError("SymbolDetector supports the QPSK / 16-QAM / 64-QAM raster constellations")
        self._qam = QAMModulator(self._mod)

    def detect_batch(self, received_symbols):
        eng = self._qam._engine()
        s = be.as_complex_tensor(received_symbols).reshape(1, -1)
        _, idx = eng.demap_count(s, want_idx=True)
        return self.constellation[be.to_numpy(idx.reshape(-1)).astype(np.int64)]

    def detect(self, received_symbol):
        return self.detect_batch(np.array([received_symbol]))[0]

    def calculate_error_rate(self, transmitted_symbols, received_symbols):
        return np.sum(np.asarray(transmitted_symbols) != np.asarray(received_symbols)) / len(transmitted_symbols)
