"""`OFDMSystem`: the PAPR / CCDF front of the reference (core/ofdm_system.py) on the GPU engine.

Mirrors the constructor, `calculate_papr*`, `transmit`, `receive`, `simulate`,
`collect_papr_for_all_modulations` and the statistics helpers.  The PAPR methods run the
`lte_papr_symbols` kernel; `collect_papr_for_all_modulations` either replays the reference's
loop draw-for-draw (`rng='numpy'`, default: `np.random.randint` bits, one transmit per
simulation) or runs the fused sweep of `lte_b200.papr.papr_sweep` (`rng='philox'`), which never
writes the time-domain stream.
"""
import time

import numpy as np

from config import LTEConfig
from core import _backend as be
from core.channel import ChannelSimulator
from core.demodulator import OFDMDemodulator, SymbolDetector
from core.modulator import OFDMModulator

_BITS_PER_SYMBOL = {'QPSK': 2, '16-QAM': 4, '64-QAM': 6}


def _stats_dict(papr, extra=None):
    out = dict(extra or {})
    out.update({'papr_mean': np.mean(papr), 'papr_max': np.max(papr), 'papr_min': np.min(papr),
                'papr_std': np.std(papr), 'num_symbols': len(papr)})
    return out


class OFDMSystem:
    def __init__(self, config=None, channel_type='awgn', itu_profile=None, frequency_ghz=2.0, velocity_kmh=0,
                 mode='lte', enable_equalization=True, enable_sc_fdm=False, rng='numpy', seed=0):
        if config is None:
            config = LTEConfig()
        self.config = config
        self.channel_type = channel_type
        self.itu_profile = itu_profile
        self.frequency_ghz = frequency_ghz
        self.velocity_kmh = velocity_kmh
        self.mode = mode
        self.enable_equalization = enable_equalization
        self.enable_sc_fdm = enable_sc_fdm
        self.rng = rng
        self.seed = seed
        self._build_chain(enable_sc_fdm)
        fs = config.fs if hasattr(config, 'fs') else 15.36e6
        if channel_type == 'rayleigh_mp':
            self.channel = ChannelSimulator(channel_type='rayleigh_mp', snr_db=10.0, fs=fs, itu_profile=itu_profile,
                                            frequency_ghz=frequency_ghz, velocity_kmh=velocity_kmh)
        else:
            self.channel = ChannelSimulator('awgn', snr_db=10.0)
        self.statistics = {'transmitted_bits': 0, 'received_bits': 0, 'bit_errors': 0, 'symbol_errors': 0,
                           'total_symbols': 0, 'transmission_time': 0}
        self.papr_values_ofdm = []
        self.papr_values_sc_fdm = []

    def _build_chain(self, sc_fdm):
        self.modulator = OFDMModulator(self.config, mode=self.mode, enable_sc_fdm=sc_fdm)
        self.demodulator = OFDMDemodulator(self.config, mode=self.mode, enable_equalization=self.enable_equalization,
                                           enable_sc_fdm=sc_fdm)
        self.symbol_detector = SymbolDetector(self.modulator.get_qam_modulator().get_constellation())

    # ------------------------------------------------------------------ PAPR
    def calculate_papr(self, signal):
        """Whole-stream PAPR (reference :80-114)."""
        x = be.as_complex_tensor(signal).reshape(-1)
        pw = x.real.double() ** 2 + x.imag.double() ** 2
        peak, avg = float(pw.max().item()), float(pw.mean().item())
        lin = peak / avg if avg > 0 else 0
        return {'papr_db': 10 * np.log10(lin) if avg > 0 else 0, 'papr_linear': lin, 'peak_power': peak,
                'avg_power': avg}

    def _per_symbol(self, signal, include_cp):
        eng = self.modulator._engine()
        x = be.as_complex_tensor(signal).reshape(1, -1)
        db, pm = eng.papr_symbols(x, include_cp=include_cp, want_peak_mean=True)
        return be.to_numpy(db.reshape(-1), np.float64), be.to_numpy(pm.reshape(-1, 2), np.float64)

    def calculate_papr_per_symbol(self, signal):
        """Per OFDM symbol including the CP (reference :116-171)."""
        db, pm = self._per_symbol(signal, True)
        return _stats_dict(db, {'papr_per_symbol': db, 'power_peak_per_symbol': pm[:, 0],
                                'power_avg_per_symbol': pm[:, 1]})

    def calculate_papr_without_cp(self, signal):
        """Per OFDM symbol over the N useful samples (reference :173-229)."""
        db, _ = self._per_symbol(signal, False)
        out = _stats_dict(db, {'papr_per_symbol': db})
        out['papr_values'] = db.tolist()
        return out

    # ------------------------------------------------------------------ link
    def transmit(self, bits, snr_db=10.0, return_time=False, use_vectorized=False):
        """bits -> modulate -> PAPR -> channel -> receive (reference :231-341); same result keys."""
        start = time.time()
        bits = np.asarray(bits)
        n = len(bits)
        self.channel.set_snr(snr_db)
        signal_tx, symbols_tx, _ = self.modulator.modulate_stream(bits)
        papr_no_cp = self.calculate_papr_without_cp(signal_tx)
        (self.papr_values_sc_fdm if self.enable_sc_fdm else self.papr_values_ofdm).extend(papr_no_cp['papr_values'])
        signal_rx, _ = self.channel.channel.transmit(signal_tx)
        bits_rx, symbols_detected, symbols_rx = self.receive(signal_rx)
        elapsed = time.time() - start
        self.statistics['transmission_time'] = elapsed
        bits_rx = np.asarray(bits_rx)
        bits_rx = np.pad(bits_rx, (0, n - len(bits_rx)), 'constant') if len(bits_rx) < n else bits_rx[:n]
        bit_errors = np.sum(bits != bits_rx)
        ber = bit_errors / n if n > 0 else 0
        flat_tx = np.concatenate(symbols_tx) if symbols_tx else np.array([])
        ncmp = min(len(flat_tx), len(symbols_detected))
        sym_err = np.sum(flat_tx[:ncmp] != symbols_detected[:ncmp]) if ncmp > 0 else 0
        ser = sym_err / ncmp if ncmp > 0 else 0
        return {
            'snr_db': snr_db, 'n_bits': n, 'transmitted_bits': n, 'received_bits': n, 'bit_errors': bit_errors,
            'errors': bit_errors, 'ber': ber, 'symbol_errors': sym_err, 'ser': ser, 'transmission_time': elapsed,
            'evm': 0.0, 'papr_no_cp': papr_no_cp, 'SNR_dB': snr_db, 'BER': ber, 'SER': ser,
            'total_symbols': len(flat_tx), 'signal_tx': signal_tx, 'tx_signal': signal_tx, 'signal_rx': signal_rx,
            'symbols_tx': flat_tx[:ncmp] if ncmp > 0 else np.array([]),
            'symbols_rx': symbols_detected[:ncmp] if ncmp > 0 else np.array([]),
            'transmitted_symbols': flat_tx, 'received_symbols': symbols_rx, 'bits_tx': bits, 'bits_rx': bits_rx,
        }

    def receive(self, signal_received):
        """-> (bits, detected constellation points, demodulated symbols) (reference :343-370)."""
        L = self.config.N + self.config.cp_length
        num = int(np.ceil(len(signal_received) / L))
        symbols, bits = self.demodulator.demodulate_stream(signal_received, num)
        return bits, self.symbol_detector.detect_batch(symbols), symbols

    def simulate(self, bits, snr_db):
        return self.transmit(bits, snr_db)

    # ------------------------------------------------------------------ CCDF collection
    def collect_papr_for_all_modulations(self, num_bits, n_simulations, snr_db=25.0, progress_callback=None):
        """PAPR samples for {QPSK, 16-QAM} x {OFDM, SC-FDM} (reference :648-735)."""
        original = (self.config.modulation, self.config.bits_per_symbol, self.enable_sc_fdm)
        results = {}
        k = 0
        try:
            for modulation in ['QPSK', '16-QAM']:
                self.config.modulation = modulation
                self.config.bits_per_symbol = self._get_bits_per_symbol(modulation)
                for sc_fdm in [False, True]:
                    k += 1
                    label = f"{modulation}_{'SC-FDM' if sc_fdm else 'OFDM'}"
                    self.enable_sc_fdm = sc_fdm
                    self._build_chain(sc_fdm)
                    if self.rng == 'philox':
                        vals = self._collect_fused(num_bits, n_simulations, sc_fdm, k)
                        if progress_callback:
                            progress_callback(int(k / 4 * 100), f"PAPR: {label}")
                    else:
                        vals = []
                        for sim in range(n_simulations):
                            bits = np.random.randint(0, 2, num_bits)
                            vals.extend(self.transmit(bits, snr_db=snr_db)['papr_no_cp']['papr_values'])
                            if progress_callback:
                                pct = ((k - 1) * n_simulations + sim + 1) / (4 * n_simulations) * 100
                                progress_callback(int(pct), f"PAPR: {label} - Sim {sim + 1}/{n_simulations}")
                    results[label] = np.array(vals) if len(vals) else np.array([])
        finally:
            self.config.modulation, self.config.bits_per_symbol, self.enable_sc_fdm = original
            self._build_chain(self.enable_sc_fdm)
        return results

    def _collect_fused(self, num_bits, n_simulations, sc_fdm, salt):
        from lte_b200.papr import papr_sweep
        eng = self.modulator._engine()
        S = eng.symbols_for_bits(num_bits)
        out = papr_sweep(eng, n_simulations, symbols_per_stream=S, sc_fdm=sc_fdm, seed=self.seed + 7919 * salt,
                         return_values=True)
        return be.to_numpy(out['values'], np.float64)

    # ------------------------------------------------------------------ helpers
    def _get_bits_per_symbol(self, modulation):
        return _BITS_PER_SYMBOL.get(modulation, 2)

    def calculate_signal_power(self, signal):
        signal = np.asarray(signal)
        return np.mean(np.abs(signal) ** 2) if np.iscomplexobj(signal) else np.mean(signal ** 2)

    def calculate_transmission_metrics(self, num_bits):
        bits_per_ofdm = self.config.Nc * self.config.bits_per_symbol
        n_sym = int(np.ceil(num_bits / bits_per_ofdm))
        duration = n_sym * (self.config.N + self.config.cp_length) * self.config.Ts
        return {'n_ofdm_symbols': n_sym, 'duration_seconds': duration,
                'throughput_mbps': (num_bits / duration) / 1e6 if duration > 0 else 0}

    def get_statistics(self):
        return self.statistics.copy()

    def reset_statistics(self):
        for key in self.statistics:
            self.statistics[key] = 0

    def get_config_info(self):
        return self.config.get_info() if hasattr(self.config, 'get_info') else {}
