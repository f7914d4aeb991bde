"""QAM mapping and OFDM modulation (reference core/modulator.py:15-420).

QAMModulator keeps the reference's natural-binary raster constellation (NOT 3GPP Gray):
the b bits of a symbol, read MSB first, index a table built with the real level in the
outer loop and the imaginary level in the inner loop (reference core/modulator.py:28-59,80-86).
OFDMModulator.modulate_stream launches the fused QAM-map + grid + IFFT + CP kernel.
"""
import numpy as np
import torch

from . import _backend as be
from .dft_precoding import SC_FDMPrecodifier
from .resource_mapper import ResourceMapper

_BITS = {'QPSK': 2, '16-QAM': 4, '64-QAM': 6}


class _QamConfig:
    """Smallest numerology that carries a plan for stand-alone map / demap calls."""
    N, Nc, cp_length, fs = 128, 76, 9, 1.92e6

    def __init__(self, bits_per_symbol):
        self.bits_per_symbol = bits_per_symbol


class QAMModulator:
    def __init__(self, modulation_type='QPSK'):
        self.modulation_type = modulation_type
        self.constellation = self._generate_constellation()
        self._bps = _BITS[modulation_type]

    def _generate_constellation(self):
        if self.modulation_type == 'QPSK':
            return np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
        if self.modulation_type == '16-QAM':
            lv, nrm = np.array([-3, -1, 1, 3]), np.sqrt(10)
        elif self.modulation_type == '64-QAM':
            lv, nrm = np.array([-7, -5, -3, -1, 1, 3, 5, 7]), np.sqrt(42)
        else:
            raise ValueError(f"Modulación no soportada: {self.modulation_type}")
        return (lv[:, None] + 1j * lv[None, :]).reshape(-1) / nrm

    def _engine(self):
        return be.engine_for(_QamConfig(self._bps), mode='simple')

    def bits_to_symbols(self, bits):
        """bits (0/1) -> complex symbols; zero-padded to a multiple of b (reference :61-88)."""
        eng = self._engine()
        b = be.as_bits_tensor(bits)
        nbits = b.shape[1]
        nsym = -(-nbits // self._bps)
        S = max(1, -(-nsym // eng.Nd))
        idx = eng.bits_to_indices(b, nbits, S)
        _, qam, _ = eng.modulate(S, idx=idx, want_qam=True, want_stats=False)
        out = qam.reshape(-1)[:nsym]
        return out if isinstance(bits, torch.Tensor) else be.to_numpy(out)

    def symbols_to_bits(self, symbols):
        """Hard decision: nearest constellation point, first minimum wins (reference :90-112)."""
        eng = self._engine()
        s = be.as_complex_tensor(symbols).reshape(1, -1)
        if s.shape[1] == 0:
            return np.array([], dtype=np.int64)
        _, idx = eng.demap_count(s, want_idx=True)
        bits = eng.indices_to_bits(idx, s.shape[1] * self._bps).reshape(-1)
        return bits if isinstance(symbols, torch.Tensor) else be.to_numpy(bits, np.int64)

    def get_constellation(self):
        return self.constellation


class OFDMModulator:
    """modes: 'simple' (symbols on bins 0..Nc-1), 'lte' (resource grid), 'sc-fdm'
    (grid + DFT precoding); reference core/modulator.py:119-302."""

    def __init__(self, config, mode='lte', enable_sc_fdm=False):
        self.config = config
        self.mode = 'sc-fdm' if enable_sc_fdm else mode
        self.enable_sc_fdm = enable_sc_fdm or (mode == 'sc-fdm')
        self.qam_modulator = QAMModulator(config.modulation)
        self.resource_mapper = ResourceMapper(config) if self.mode in ['lte', 'sc-fdm'] else None
        if self.enable_sc_fdm and self.resource_mapper is not None:
            self.sc_fdm_precoder = SC_FDMPrecodifier(
                num_data_subcarriers=len(self.resource_mapper.get_data_indices()), enable=self.enable_sc_fdm)
        else:
            self.sc_fdm_precoder = None
        self._engine_mode = 'lte' if self.resource_mapper is not None else 'simple'

    def _engine(self):
        return be.engine_for(self.config, mode=self._engine_mode)

    # -- device-level core shared with OFDMSimulator ------------------------------------
    def _modulate_stream_device(self, bits_t, num_ofdm_symbols=None):
        """bits_t: uint8 CUDA [1, n] -> (tx [1, S*L], qam [S, Nd], idx [1, S*Nd], stats, S)."""
        eng = self._engine()
        nbits = bits_t.shape[1]
        S = eng.symbols_for_bits(nbits) if num_ofdm_symbols is None else int(num_ofdm_symbols)
        idx = eng.bits_to_indices(bits_t, nbits, S)
        if self.enable_sc_fdm and self.sc_fdm_precoder is not None:
            qam = eng.qam_map(idx)
            pre = eng.dft_m(qam.reshape(S, eng.Nd), eng.Nd)
            tx, _, stats = eng.modulate(S, symbols=pre.reshape(1, -1))
        else:
            tx, qam, stats = eng.modulate(S, idx=idx, want_qam=True)
        return tx, qam.reshape(S, eng.Nd), idx, stats, S

    def modulate(self, bits):
        """One OFDM symbol: (signal[N+cp], qam_symbols, mapping_info) -- reference :158-250."""
        eng = self._engine()
        b = be.as_bits_tensor(bits)
        b = b[:, :eng.Nd * self.config.bits_per_symbol]       # extra symbols are dropped (:228-230)
        tx, qam, _, _, _ = self._modulate_stream_device(b, num_ofdm_symbols=1)
        nsym = min(-(-b.shape[1] // self.config.bits_per_symbol), eng.Nd)
        qam_np = be.to_numpy(qam.reshape(-1))
        if self._engine_mode == 'lte':
            return be.to_numpy(tx.reshape(-1)), qam_np, self._mapping_info(eng)
        return be.to_numpy(tx.reshape(-1)), qam_np[:nsym], None

    def _mapping_info(self, eng):
        g = self.resource_mapper.grid
        return {'num_data_mapped': eng.Nd, 'num_pilots_mapped': eng.Np,
                'num_nulls': len(g.get_guard_indices()) + 1, 'data_indices': g.get_data_indices(),
                'pilot_indices': g.get_pilot_indices(), 'guard_indices': g.get_guard_indices(),
                'dc_index': g.dc_index, 'grid_statistics': self.resource_mapper.stats}

    def modulate_stream(self, bits, num_ofdm_symbols=None):
        """(signal_concatenated, [symbols per OFDM symbol], mapping_infos) -- reference :252-302."""
        b = be.as_bits_tensor(bits)
        tx, qam, _, _, S = self._modulate_stream_device(b, num_ofdm_symbols)
        if self._engine_mode == 'lte':
            be.reference_pilot_side_effect(self.resource_mapper.pilot_pattern.cell_id, self._engine().Np)
        qam_np = be.to_numpy(qam)
        infos = [self._mapping_info(self._engine()) for _ in range(S)] if self._engine_mode == 'lte' else None
        return be.to_numpy(tx.reshape(-1)), [qam_np[s] for s in range(S)], infos

    modulate_stream_vectorized = modulate_stream

    def set_sc_fdm_enabled(self, enable):
        """Switch the DFT precoder on/off (reference :405-416)."""
        self.enable_sc_fdm = enable
        self.mode = 'sc-fdm' if enable else 'lte'
        if self.sc_fdm_precoder is not None:
            self.sc_fdm_precoder.set_enable(enable)

    def get_qam_modulator(self):
        return self.qam_modulator


# ------------------------------------------------------------------ soft demapping (reference :423-540)
class _OneBinConfig:
    """simple-mode plan with a single data position: every symbol is its own stream, so each one can carry
    its own noise variance through lte_soft_demap's per-stream sigma2."""
    N, Nc, cp_length, fs = 64, 1, 0, 1.92e6

    def __init__(self, bits_per_symbol):
        self.bits_per_symbol = bits_per_symbol


def symbols_to_llrs(symbols, noise_var, bits_per_symbol):
    """Max-log LLRs of `symbols` with a scalar or per-symbol noise variance on `lte_soft_demap`
    (QPSK: exact, unclipped; 16/64-QAM: clipped to +-10) -- the arithmetic of
    OFDMSimulator._calculate_llrs_qpsk/_16qam/_64qam (reference core/ofdm_core.py:791-923)."""
    import torch
    y = np.asarray(symbols)
    if y.size == 0:
        return np.array([], dtype=np.float64)
    eng = be.engine_for(_OneBinConfig(bits_per_symbol), mode='simple')
    nv = np.broadcast_to(np.asarray(noise_var, dtype=np.float32), y.shape)
    data = be.as_complex_tensor(y.reshape(-1, 1))
    s2 = torch.from_numpy(np.array(nv.reshape(-1), dtype=np.float32)).to(data.device)
    llr = eng.soft_demap(data, None, s2, False, 1, 1)
    return be.to_numpy(llr.reshape(-1)).astype(np.float64)


def qpsk_to_llrs(symbols, noise_var):
    """[LLR_I0, LLR_Q0, LLR_I1, ...] = (2 / noise_var) * Re/Im(y) * sqrt(2) (reference :423-475)."""
    return symbols_to_llrs(symbols, noise_var, 2)


def qam16_to_llrs(symbols, noise_var):
    """The reference leaves this module function unimplemented (:478-500) and does the 16-QAM LLRs in
    OFDMSimulator._calculate_llrs_16qam; the error is kept so callers that probe for it see the same thing."""
    raise NotImplementedError("16-QAM LLR generation not yet implemented")


def qam64_to_llrs(symbols, noise_var):
    raise NotImplementedError("64-QAM LLR generation not yet implemented")
