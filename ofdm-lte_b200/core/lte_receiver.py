"""LTE receiver: FFT demodulation, CRS least-squares channel estimation with linear frequency
interpolation, zero-forcing equalisation and hard detection (reference core/lte_receiver.py).
Every numerical step is a launch of the CUDA stage kernels through `lte_b200.LinkEngine`."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from lte_b200 import _native as nat

from . import _backend as be
from .modulator import QAMModulator
from .resource_mapper import LTEResourceGrid, PilotPattern


class LTEChannelEstimator:
    """reference core/lte_receiver.py:20-133."""

    def __init__(self, config, cell_id=0):
        self.config = config
        self.cell_id = cell_id
        self.resource_grid = LTEResourceGrid(config.N, config.Nc)
        self.pilot_pattern = PilotPattern(cell_id)

    def _engine(self):
        if self.cell_id == 0:
            return be.engine_for(self.config)
        from lte_b200 import tables
        return be.engine_for(self.config, pilot_sets=tables.pilot_values(
            self.cell_id, len(self.resource_grid.get_pilot_indices()))[None, :])

    def estimate_channel(self, received_signal, tx_signal: Optional[np.ndarray] = None) -> Dict:
        """received_signal: one OFDM symbol in the frequency domain, length N."""
        eng = self._engine()
        pilot_indices = self.resource_grid.get_pilot_indices()
        known = self.pilot_pattern.generate_pilots(len(pilot_indices))        # keeps the RNG side effect
        Y = be.as_complex_tensor(received_signal).reshape(1, 1, self.config.N)
        H = eng.estimate(Y, 1, 1, nat.WINDOW_FULL).reshape(-1)
        pidx = torch.from_numpy(pilot_indices).to(H.device)
        rp = Y.reshape(-1)[pidx]
        kp = be.as_complex_tensor(known)
        pilot_power = torch.mean(rp.abs() ** 2)
        noise_power = torch.mean((rp - kp).abs() ** 2)
        pilot_snr = float(pilot_power / (noise_power + 1e-10))
        return {
            'channel_estimate': be.to_numpy(H),
            'pilot_channel': be.to_numpy(H[pidx]),
            'pilot_indices': pilot_indices,
            'pilot_snr_linear': pilot_snr,
            'pilot_snr_db': 10 * np.log10(pilot_snr + 1e-10),
            'interpolated': True,
        }


class LTEEqualizerZF:
    """Y / (H + 1e-6) on every bin (reference core/lte_receiver.py:136-180)."""

    def __init__(self, config, regularization=1e-6):
        self.config = config
        self.regularization = regularization

    def equalize(self, received_symbols, channel_estimate):
        y = be.as_complex_tensor(received_symbols)
        h = be.as_complex_tensor(channel_estimate)
        out = y / (h + self.regularization)
        return out if isinstance(received_symbols, torch.Tensor) else be.to_numpy(out)


class LTEReceiver:
    """reference core/lte_receiver.py:183-557."""

    def __init__(self, config, cell_id=0, enable_equalization=True, enable_sc_fdm=False):
        self.config = config
        self.cell_id = cell_id
        self.enable_equalization = enable_equalization
        self.enable_sc_fdm = enable_sc_fdm
        self.resource_grid = LTEResourceGrid(config.N, config.Nc)
        self.pilot_pattern = PilotPattern(cell_id)
        self.channel_estimator = LTEChannelEstimator(config, cell_id)
        self.equalizer = LTEEqualizerZF(config)
        self.qam_demodulator = QAMModulator(config.modulation)
        if enable_sc_fdm:
            from .dft_precoding import SC_FDMDecodifier
            self.sc_fdm_decoder = SC_FDMDecodifier(len(self.resource_grid.get_data_indices()), enable=True)
        else:
            self.sc_fdm_decoder = None
        self.channel_estimates = []
        self.equalization_info = []
        self.slot_size = nat.LTE_SLOT_SYMBOLS
        self.faithful_rng = True     # reproduce the reference's global-RNG re-seeding on estimation

    def _engine(self):
        return self.channel_estimator._engine()

    # -- device-level core ----------------------------------------------------------------
    def _num_symbols(self, n_samples):
        return max(n_samples // (self.config.N + self.config.cp_length), 1)

    def _fft_device(self, rx_t):
        """rx_t: complex64 CUDA [rows, n] -> (Y [rows, S, N], S); short streams are zero-padded
        to one symbol, trailing partial symbols dropped (reference :459-481)."""
        eng = self._engine()
        rows, n = rx_t.shape
        S = self._num_symbols(n)
        need = S * eng.L
        if n < need:
            rx_t = torch.nn.functional.pad(rx_t, (0, need - n))
        elif n > need:
            rx_t = rx_t[:, :need].contiguous()
        return eng.rx_fft(rx_t, rows, S, nat.WINDOW_FULL), S

    def _receive_device(self, rx_t):
        """Full SISO receive chain on the device.  Returns dict of CUDA tensors."""
        eng = self._engine()
        Y, S = self._fft_device(rx_t)
        H = eng.estimate(Y, 1, S, nat.WINDOW_FULL)
        if self.faithful_rng:
            be.reference_pilot_side_effect(self.cell_id, eng.Np)
        data = eng.zf(Y, H if self.enable_equalization else None, 1, S, nat.WINDOW_FULL)
        if self.enable_sc_fdm and self.sc_fdm_decoder is not None:
            data = eng.dft_m(data.reshape(S, eng.Nd), eng.Nd, inverse=True).reshape(1, -1)
        return dict(Y=Y, H=H, data=data, S=S)

    # -- reference-shaped methods ---------------------------------------------------------------
    def receive_and_decode(self, received_ofdm_signal) -> Dict:
        rx = be.as_complex_tensor(received_ofdm_signal).reshape(1, -1)
        eng = self._engine()
        if rx.shape[1] == 0:
            e = np.array([])
            return {'symbols_received': e, 'symbols_equalized': e, 'symbols_data_only': e, 'symbols_detected': e,
                    'bits': e, 'channel_estimate': e, 'channel_snr_db': 0, 'pilot_snr_db': 0,
                    'num_data_symbols': 0, 'num_pilot_symbols': 0, 'equalization_enabled': self.enable_equalization}
        r = self._receive_device(rx)
        Y, H, data, S = r['Y'], r['H'], r['data'], r['S']
        Hs = H.reshape(-1, eng.N).repeat_interleave(self.slot_size, dim=0)[:S]
        eq = (Y.reshape(S, eng.N) / (Hs + 1e-6)) if self.enable_equalization else Y.reshape(S, eng.N)
        _, idx = eng.demap_count(data, want_idx=True)
        bits = eng.indices_to_bits(idx, idx.shape[1] * eng.bps).reshape(-1)
        const = be.as_complex_tensor(self.qam_demodulator.constellation)
        detected = const[idx.reshape(-1).long()]
        snr_db = self._pilot_snr_db(Y[0], S)
        ch0 = be.to_numpy(H.reshape(-1, eng.N)[0])
        self.channel_estimates.append(ch0)
        self.equalization_info.append({'channel_snr_db': snr_db, 'num_data_symbols': data.shape[1]})
        return {
            'symbols_received': be.to_numpy(Y.reshape(-1)),
            'symbols_equalized': be.to_numpy(eq.reshape(-1)),
            'symbols_data_only': be.to_numpy(data.reshape(-1)),
            'symbols_detected': be.to_numpy(detected),
            'bits': be.to_numpy(bits, np.int64),
            'channel_estimate': ch0,
            'channel_snr_db': snr_db,
            'pilot_snr_db': snr_db,
            'num_data_symbols': data.shape[1],
            'num_pilot_symbols': eng.Np * S,
            'equalization_enabled': self.enable_equalization,
        }

    def _pilot_snr_db(self, Y, S):
        """Mean over the slots of the pilot SNR statistic of their first symbol (reference :78-80, :393, :409)."""
        pidx = torch.from_numpy(self.resource_grid.get_pilot_indices()).to(Y.device)
        kp = be.as_complex_tensor(self.pilot_pattern.generate_pilots(len(pidx)))
        snrs = []
        for s0 in range(0, S, self.slot_size):
            rp = Y[s0][pidx]
            snr = torch.mean(rp.abs() ** 2) / (torch.mean((rp - kp).abs() ** 2) + 1e-10)
            snrs.append(10 * np.log10(float(snr) + 1e-10))
        return float(np.mean(snrs)) if snrs else 0.0

    def _estimate_channel_periodic(self, all_received_symbols: List[np.ndarray]) -> Tuple[List[np.ndarray], float]:
        """One LS estimate per 14-symbol slot, held for the slot (reference :360-411)."""
        eng = self._engine()
        S = len(all_received_symbols)
        Y = be.as_complex_tensor(np.stack([np.asarray(a) for a in all_received_symbols])).reshape(1, S, eng.N)
        H = eng.estimate(Y, 1, S, nat.WINDOW_FULL).reshape(-1, eng.N)
        be.reference_pilot_side_effect(self.cell_id, eng.Np)
        Hn = be.to_numpy(H)
        return [Hn[s // self.slot_size] for s in range(S)], self._pilot_snr_db(Y[0], S)

    def _demodulate_ofdm_stream(self, received_signal) -> List[np.ndarray]:
        rx = be.as_complex_tensor(received_signal).reshape(1, -1)
        Y, S = self._fft_device(rx)
        Yn = be.to_numpy(Y.reshape(S, -1))
        return [Yn[s] for s in range(S)]

    def _demodulate_ofdm(self, received_signal):
        syms = self._demodulate_ofdm_stream(received_signal)
        return np.concatenate(syms) if len(syms) > 1 else (syms[0] if syms else np.array([]))

    def _detect_symbols(self, received_symbols):
        eng = self._engine()
        s = be.as_complex_tensor(received_symbols).reshape(1, -1)
        _, idx = eng.demap_count(s, want_idx=True)
        const = be.as_complex_tensor(self.qam_demodulator.constellation)
        return be.to_numpy(const[idx.reshape(-1).long()])

    def calculate_ber(self, transmitted_bits, received_bits) -> Dict:
        n = min(len(transmitted_bits), len(received_bits))
        errors = int(np.sum(np.asarray(transmitted_bits[:n]) != np.asarray(received_bits[:n])))
        return {'ber': errors / n if n > 0 else 0, 'errors': errors, 'total_bits': n}

    def reset_history(self):
        self.channel_estimates = []
        self.equalization_info = []

    def get_channel_estimate_history(self):
        return np.array(self.channel_estimates)
