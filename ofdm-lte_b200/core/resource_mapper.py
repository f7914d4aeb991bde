"""LTE resource grid, pilot pattern and symbol mapping.

Same classes and methods as the reference's core/resource_mapper.py (LTEResourceGrid :17-111,
PilotPattern :114-152, ResourceMapper :155-266).  The bin classification is integer table work
done once per grid; `map_symbols` places symbols with device-side indexing.  In the link chain
itself the mapping is fused into the IFFT loader (csrc/ofdm.cu, tx_map_ifft_kernel).
"""
from typing import Dict, Tuple

import numpy as np
import torch

from lte_b200 import tables

from . import _backend as be


class LTEResourceGrid:
    """Bin classes in raw FFT-bin order: guards [0, gl) and [N-gr, N), DC at N/2, pilots where
    (k - gl) % 6 == 3, data elsewhere (reference core/resource_mapper.py:45-74)."""

    def __init__(self, N: int, Nc: int):
        self.N, self.Nc = N, Nc
        self.num_guard_left = (N - Nc) // 2
        self.num_guard_right = N - Nc - self.num_guard_left
        self.dc_index = N // 2
        self.pilot_spacing = 6
        self._init_subcarrier_types()

    def _init_subcarrier_types(self):
        self._data, self._pilot = tables.grid_indices(self.N, self.Nc)
        k = np.arange(self.N)
        self._guard = k[(k < self.num_guard_left) | (k >= self.N - self.num_guard_right)]
        kinds = np.full(self.N, 'guard', dtype=object)
        kinds[self._data] = 'data'
        kinds[self._pilot] = 'pilot'
        if self.num_guard_left <= self.dc_index < self.N - self.num_guard_right:
            kinds[self.dc_index] = 'dc'
        self.subcarrier_types = {int(i): kinds[i] for i in range(self.N)}

    def get_subcarrier_type(self, k: int) -> str:
        return self.subcarrier_types.get(k, 'guard')

    def get_data_indices(self) -> np.ndarray:
        return self._data.copy()

    def get_pilot_indices(self) -> np.ndarray:
        return self._pilot.copy()

    def get_guard_indices(self) -> np.ndarray:
        return self._guard.copy()

    def get_statistics(self) -> Dict:
        return {
            'total_subcarriers': self.N,
            'useful_subcarriers': self.Nc,
            'data_subcarriers': len(self._data),
            'pilot_subcarriers': len(self._pilot),
            'guard_subcarriers': len(self._guard),
            'dc_subcarriers': 1,
            'guard_left': self.num_guard_left,
            'guard_right': self.num_guard_right,
            'pilot_spacing': self.pilot_spacing,
        }


class PilotPattern:
    """(+-1)(1+1j)/sqrt(2) pilots from NumPy's legacy stream seeded by cell_id
    (reference core/resource_mapper.py:114-152)."""

    faithful_global_rng = True   # reproduce the reference's np.random.seed(cell_id) side effect

    def __init__(self, cell_id: int = 0, pilot_symbol_value: complex = None):
        self.cell_id = cell_id
        self.pilot_symbol_value = (1 + 1j) / np.sqrt(2) if pilot_symbol_value is None else pilot_symbol_value

    def generate_pilots(self, num_pilots: int) -> np.ndarray:
        if self.faithful_global_rng:
            np.random.seed(self.cell_id)
            phases = np.random.choice([1, -1], size=num_pilots)
        else:
            phases = np.random.RandomState(self.cell_id).choice([1, -1], size=num_pilots)
        return self.pilot_symbol_value * phases


class ResourceMapper:
    """Places data and pilot symbols on one OFDM symbol's grid
    (reference core/resource_mapper.py:155-246)."""

    def __init__(self, config, cell_id: int = 0):
        self.config = config
        self.grid = LTEResourceGrid(config.N, config.Nc)
        self.pilot_pattern = PilotPattern(cell_id)
        self.stats = self.grid.get_statistics()

    def map_symbols(self, data_symbols) -> Tuple[np.ndarray, Dict]:
        data_indices = self.grid.get_data_indices()
        pilot_indices = self.grid.get_pilot_indices()
        pilots = self.pilot_pattern.generate_pilots(len(pilot_indices))
        n = min(len(data_symbols), len(data_indices))
        dev = be.device()
        grid = torch.zeros(self.config.N, dtype=torch.complex64, device=dev)
        grid[torch.from_numpy(data_indices[:n]).to(dev)] = be.as_complex_tensor(data_symbols).reshape(-1)[:n]
        grid[torch.from_numpy(pilot_indices).to(dev)] = be.as_complex_tensor(pilots)
        mapping_info = {
            'num_data_mapped': n,
            'num_pilots_mapped': len(pilot_indices),
            'num_nulls': len(self.grid.get_guard_indices()) + 1,
            'data_indices': data_indices[:n],
            'pilot_indices': pilot_indices,
            'guard_indices': self.grid.get_guard_indices(),
            'dc_index': self.grid.dc_index,
            'grid_statistics': self.stats,
        }
        return be.to_numpy(grid), mapping_info

    def extract_pilots(self, received_grid) -> Tuple[np.ndarray, np.ndarray]:
        pilot_indices = self.grid.get_pilot_indices()
        return pilot_indices, np.asarray(received_grid)[pilot_indices]

    def get_data_indices(self) -> np.ndarray:
        return self.grid.get_data_indices()

    def get_statistics(self) -> Dict:
        return self.stats
