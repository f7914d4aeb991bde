"""Legacy 2 x N estimator: one full-pilot LTEChannelEstimator (own cell id) per TX antenna
(reference core/mimo_channel_estimator.py:8-125).  Not used by the orchestrators; kept for the
API surface named in SURVEY 8b.  Each estimate is a `lte_crs_ls_interp` launch."""
import itertools

import numpy as np

from .lte_receiver import LTEChannelEstimator


class MIMOChannelEstimator:
    def __init__(self, config, num_tx=2, num_rx=1):
        self.config, self.num_tx, self.num_rx = config, num_tx, num_rx
        self.estimators = [LTEChannelEstimator(config, cell_id=tx) for tx in range(num_tx)]

    def estimate_mimo_channel(self, received_grids: list, transmitted_pilots: list = None) -> dict:
        """H[r, t, :] = estimate of RX grid r with the pilots of TX t (all N bins, plus the data-bin view)."""
        H = np.zeros((self.num_rx, self.num_tx, self.config.N), dtype=np.complex64)
        snrs = []
        for r, t in itertools.product(range(self.num_rx), range(self.num_tx)):
            est = self.estimators[t].estimate_channel(received_grids[r])
            H[r, t] = est['channel_estimate']
            snrs.append(est['pilot_snr_db'])
        data_bins = self.estimators[0].resource_grid.get_data_indices()
        return dict(channel_matrix=H, channel_matrix_data=H[:, :, data_bins], data_indices=data_bins,
                    snr_db=np.mean(snrs) if snrs else 0, shape=H.shape)

    def extract_channel_for_alamouti(self, channel_matrix_data, rx_idx: int = 0) -> tuple:
        h = channel_matrix_data[rx_idx]
        return h[0], h[1]


def estimate_mimo_channel_simple(config, received_grids: list, num_tx: int = 2):
    est = MIMOChannelEstimator(config, num_tx=num_tx, num_rx=len(received_grids))
    return est.estimate_mimo_channel(received_grids)
