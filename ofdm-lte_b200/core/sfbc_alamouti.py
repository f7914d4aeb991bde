"""SFBC Alamouti encoder / decoder and grid mapper (reference core/sfbc_alamouti.py)."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from lte_b200 import _native as nat
from lte_b200 import tables

from . import _backend as be


class _PairConfig:
    """Plan whose data-bin count equals the number of symbols handed to encode()/decode()."""
    cp_length, fs, bits_per_symbol = 0, 1.92e6, 2

    def __init__(self, n):
        self.Nc = n
        self.N = max(64, 1 << int(np.ceil(np.log2(max(n, 1)))))


class SFBCAlamouti:
    def __init__(self, num_tx: int = 2, enabled: bool = True):
        if num_tx != 2:
            raise ValueError("Alamouti SFBC requires exactly 2 TX antennas")
        self.num_tx = num_tx
        self.enabled = enabled

    def encode(self, symbols) -> Tuple[np.ndarray, np.ndarray]:
        """TX0 [s0, -s1*], TX1 [s1, s0*] per pair (reference :45-78)."""
        if not self.enabled:
            return np.array(symbols, copy=True), np.array(symbols, copy=True)
        n = len(symbols)
        if n % 2 != 0:
            raise ValueError(f"Number of symbols must be even for Alamouti coding, got {n}")
        eng = be.engine_for(_PairConfig(n), mode='simple')
        out, _ = eng.sfbc_encode(1, symbols=be.as_complex_tensor(symbols).reshape(1, -1))
        o = be.to_numpy(out)
        return o[0, :n], o[1, :n]

    def decode(self, rx_symbols, H0, H1, regularization: float = 1e-10) -> np.ndarray:
        """reference :80-163 (regularization is the reference's fixed 1e-10)."""
        if not self.enabled:
            return np.array(rx_symbols, copy=True)
        n = len(rx_symbols)
        if n % 2 != 0:
            raise ValueError(f"Number of RX symbols must be even, got {n}")
        if len(H0) != n or len(H1) != n:
            raise ValueError(f"Channel estimates must have length {n}")
        eng = be.engine_for(_PairConfig(n), mode='simple')
        Y = be.as_complex_tensor(rx_symbols).reshape(1, 1, n)
        out = eng.sfbc_decode(Y, be.as_complex_tensor(H0).reshape(1, 1, n), be.as_complex_tensor(H1).reshape(1, 1, n),
                              1, 1, 1, nat.WINDOW_USEFUL)
        return be.to_numpy(out.reshape(-1))

    def get_statistics(self) -> Dict:
        return {'enabled': self.enabled, 'num_tx': self.num_tx, 'coding_scheme': 'Alamouti SFBC', 'rate': 1.0,
                'diversity_order': 2}


class SFBCResourceMapper:
    """reference core/sfbc_alamouti.py:176-325."""

    def __init__(self, resource_mapper):
        self.resource_mapper = resource_mapper
        self.data_indices = resource_mapper.get_data_indices()
        self.num_data = len(self.data_indices)
        if self.num_data % 2 != 0:
            self.num_data -= 1
            self.data_indices = self.data_indices[:self.num_data]

    def prepare_data_for_sfbc(self, qam_symbols):
        q = np.asarray(qam_symbols)
        if len(q) < self.num_data:
            q = np.pad(q, (0, self.num_data - len(q)), 'constant', constant_values=0)
        return q[:self.num_data]

    def map_sfbc_to_grid(self, tx0_symbols, tx1_symbols, pilot_symbols: Optional[np.ndarray] = None):
        cfg = self.resource_mapper.config
        pilot_idx = self.resource_mapper.grid.get_pilot_indices()
        sets = tables.mimo_pilot_sets(2, len(pilot_idx))
        from .resource_mapper import PilotPattern
        for cell, sl in ((0, pilot_idx[::2]), (1, pilot_idx[1::2])):   # keeps the RNG side effect (:252-256)
            PilotPattern(cell).generate_pilots(len(sl))
        dev = be.device()
        grids = torch.zeros((2, cfg.N), dtype=torch.complex64, device=dev)
        di = torch.from_numpy(self.data_indices).to(dev)
        grids[0, di] = be.as_complex_tensor(tx0_symbols).reshape(-1)[:self.num_data]
        grids[1, di] = be.as_complex_tensor(tx1_symbols).reshape(-1)[:self.num_data]
        pi = torch.from_numpy(pilot_idx).to(dev)
        ps = be.as_complex_tensor(sets)
        own = ps != 0
        for t in range(2):
            grids[t, pi[own[t]]] = ps[t][own[t]]
        g = be.to_numpy(grids)
        return g[0], g[1]

    def extract_data_from_grid(self, rx_grid):
        return np.asarray(rx_grid)[self.data_indices]

    def apply_generic_precoding(self, symbols, W_matrix) -> List[np.ndarray]:
        s = be.as_complex_tensor(symbols)
        if s.ndim == 1:
            s = s.reshape(1, -1)
        W = be.as_complex_tensor(W_matrix)
        if W.shape[1] != s.shape[0]:
            raise ValueError(f"W_matrix shape {tuple(W.shape)} no compatible con {s.shape[0]} layers")
        out = be.to_numpy(W @ s)
        return [out[i] for i in range(out.shape[0])]
