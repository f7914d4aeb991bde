"""Legacy 2 x N estimator: one full-pilot LTEChannelEstimator (own cell id) per TX antenna
(reference core/mimo_channel_estimator.py:8-125).  Not used by the orchestrators; kept for the
API surface named in SURVEY 8b.  Each estimate is a `lte_crs_ls_interp` launch."""
import numpy as np

from .lte_receiver import LTEChannelEstimator


class MIMOChannelEstimator:
    def __init__(self, config, num_tx=2, num_rx=1):
        self.config = config
        self.num_tx = num_tx
        self.num_rx = num_rx
        self.estimators = [LTEChannelEstimator(config, cell_id=t) for t in range(num_tx)]

    def estimate_mimo_channel(self, received_grids: list, transmitted_pilots: list = None) -> dict:
        N = self.config.N
        H_full = np.zeros((self.num_rx, self.num_tx, N), dtype=np.complex64)
        snr_list = []
        for r in range(self.num_rx):
            for t in range(self.num_tx):
                info = self.estimators[t].estimate_channel(received_grids[r])
                H_full[r, t, :] = info['channel_estimate']
                snr_list.append(info['pilot_snr_db'])
        data_indices = self.estimators[0].resource_grid.get_data_indices()
        return {'channel_matrix': H_full, 'channel_matrix_data': H_full[:, :, data_indices],
                'data_indices': data_indices, 'snr_db': np.mean(snr_list) if snr_list else 0,
                'shape': H_full.shape}

    def extract_channel_for_alamouti(self, channel_matrix_data, rx_idx: int = 0) -> tuple:
        return channel_matrix_data[rx_idx, 0, :], channel_matrix_data[rx_idx, 1, :]


def estimate_mimo_channel_simple(config, received_grids: list, num_tx: int = 2):
    return MIMOChannelEstimator(config, num_tx=num_tx, num_rx=len(received_grids)).estimate_mimo_channel(received_grids)
