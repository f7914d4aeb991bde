"""N-TX CRS channel estimation with frequency-interleaved pilot subsets
(reference core/mimo_channel_estimator_periodic.py).  TX t owns pilot_idx[t % step :: step],
step = min(num_tx, 4), with the sign stream of cell t % 4; every (rx, tx) estimate is one launch
batch of `lte_crs_ls_interp` with the matching pilot set."""
from typing import Dict, List, Tuple

import numpy as np
import torch

from lte_b200 import _native as nat
from lte_b200 import tables

from . import _backend as be
from .lte_receiver import LTEChannelEstimator
from .resource_mapper import LTEResourceGrid, PilotPattern


class MIMOChannelEstimatorPeriodic:
    def __init__(self, config, num_tx: int = 2, num_rx: int = 2, slot_size: int = 14):
        self.config = config
        self.num_tx = num_tx
        self.num_rx = num_rx
        self.slot_size = slot_size
        if num_tx not in [2, 4, 8]:
            raise ValueError(f"num_tx debe ser 2, 4 o 8, recibido: {num_tx}")
        self.estimators = [LTEChannelEstimator(config, cell_id=t % 4) for t in range(num_tx)]
        self.pilot_patterns = [PilotPattern(cell_id=t % 4) for t in range(num_tx)]
        self.resource_grid = LTEResourceGrid(config.N, config.Nc)
        self.faithful_rng = True

    def _engine(self):
        npil = len(self.resource_grid.get_pilot_indices())
        return be.engine_for(self.config, pilot_sets=tables.mimo_pilot_sets(self.num_tx, npil))

    def get_orthogonal_pilot_indices(self) -> List[np.ndarray]:
        allp = self.resource_grid.get_pilot_indices()
        step = self.num_tx if self.num_tx <= 4 else 4
        return [allp[t % step::step] for t in range(self.num_tx)]

    def _side_effect(self):
        if self.faithful_rng:                       # last generate_pilots call of the reference loop
            t = self.num_tx - 1
            be.reference_pilot_side_effect(t % 4, len(self.get_orthogonal_pilot_indices()[t]))

    def _estimate_device(self, Y, rows, S, every_symbol=False):
        """Y [rows, S, N] -> H [T][rows, nest, N]; nest = S when every_symbol else ceil(S/14)."""
        eng = self._engine()
        if every_symbol:
            Hs = [eng.estimate(Y.reshape(rows * S, 1, eng.N), rows * S, 1, nat.WINDOW_FULL, pilot_set=t)
                  .reshape(rows, S, eng.N) for t in range(self.num_tx)]
        else:
            Hs = [eng.estimate(Y, rows, S, nat.WINDOW_FULL, pilot_set=t) for t in range(self.num_tx)]
        self._side_effect()
        return Hs

    def estimate_channel_from_grid(self, grid_rx, return_full_freq: bool = True) -> Tuple[np.ndarray, Dict]:
        g = np.asarray(grid_rx)
        grids = g.reshape(1, -1) if g.ndim == 1 else g
        R, N = grids.shape
        Y = be.as_complex_tensor(grids).reshape(R, 1, N)
        Hs = self._estimate_device(Y, R, 1)
        H = torch.stack([h.reshape(R, N) for h in Hs], dim=1)           # [R, T, N]
        own = self.get_orthogonal_pilot_indices()
        if not return_full_freq:
            H = torch.stack([H[:, t, torch.from_numpy(own[t]).to(H.device)].mean(dim=1)
                             for t in range(self.num_tx)], dim=1)
        info = {'num_pilots_per_tx': [len(o) for o in own], 'pilot_indices': own, 'num_rx': R,
                'num_tx': self.num_tx, 'N': N}
        return be.to_numpy(H), info

    def estimate_channel_periodic(self, all_received_grids):
        """One estimate per slot, held for the slot.  (The reference's version raises at HEAD because
        it unpacks three values from estimate_channel_from_grid; this is the intended behaviour.)"""
        S = len(all_received_grids)
        Y = be.as_complex_tensor(np.stack([np.asarray(g) for g in all_received_grids])).reshape(1, S, -1)
        Hs = self._estimate_device(Y, 1, S)
        H0, H1 = be.to_numpy(Hs[0][0]), be.to_numpy(Hs[1][0])
        return ([H0[s // self.slot_size] for s in range(S)], [H1[s // self.slot_size] for s in range(S)], 0.0)

    def demodulate_and_estimate_mimo(self, signal_rx, cp_length: int):
        eng = self._engine()
        rx = be.as_complex_tensor(signal_rx).reshape(1, -1)
        S = rx.shape[1] // eng.L
        Y = eng.rx_fft(rx[:, :S * eng.L].contiguous(), 1, S, nat.WINDOW_FULL)
        Hs = self._estimate_device(Y, 1, S)
        Yn, H0, H1 = be.to_numpy(Y[0]), be.to_numpy(Hs[0][0]), be.to_numpy(Hs[1][0])
        return ([Yn[s] for s in range(S)], [H0[s // self.slot_size] for s in range(S)],
                [H1[s // self.slot_size] for s in range(S)])
