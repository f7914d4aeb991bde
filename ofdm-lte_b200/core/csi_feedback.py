"""CSI feedback (PMI / CQI / RI) for the beamforming path -- reference core/csi_feedback.py:25-228.
PMI selection, the effective channel and the post-precoding power run on the GPU
(`lte_bf_weights`, one thread per channel matrix, fp64 registers); `feedback_batch` does it for a
whole batch of channels in one launch.  CQI quantisation and the rank indicator are scalar
bookkeeping on the selected entry."""
import numpy as np
import torch

from . import _backend as be
from .codebook_lte import LTECodebook


class _BinsConfig:
    """Smallest plan that can carry the precoder kernels (they do not touch the grid tables)."""
    N, Nc, cp_length, fs, bits_per_symbol = 128, 76, 9, 1.92e6, 2


def _bf_engine():
    return be.engine_for(_BinsConfig)


class CSIFeedback:
    def __init__(self, num_tx, num_rx, codebook_type='TM6', feedback_mode='perfect'):
        self.num_tx = num_tx
        self.num_rx = num_rx
        self.codebook_type = codebook_type
        self.feedback_mode = feedback_mode
        self.codebook = LTECodebook(num_tx, transmission_mode=codebook_type)
        self.total_feedbacks = 0
        self.pmi_history = []

    # -- batched GPU entry: h complex64 CUDA tensor [B, R, T] -> (pmi int32 [B], heff [B, R])
    def feedback_batch(self, h):
        _, heff, pmi, _ = _bf_engine().bf_weights(h, self.codebook.codebook, mode='CODEBOOK')
        return pmi, heff

    def _h_device(self, H_channel):
        H = np.asarray(H_channel, dtype=np.complex64)
        if H.ndim == 1:
            H = H.reshape(1, -1)
        return be.as_complex_tensor(H[None])

    def calculate_pmi(self, H_channel):
        pmi, _ = self.feedback_batch(self._h_device(H_channel))
        pmi = int(pmi.item())
        self.pmi_history.append(pmi)
        self.total_feedbacks += 1
        return pmi

    def calculate_cqi(self, H_channel, pmi, noise_variance=1.0):
        H_eff = np.asarray(H_channel) @ self.codebook.get_precoder(pmi)
        sinr_db = 10 * np.log10(np.sum(np.abs(H_eff) ** 2) / noise_variance)
        return self._sinr_to_cqi(sinr_db), sinr_db

    @staticmethod
    def _sinr_to_cqi(sinr_db):
        """2 dB steps from -6 dB (reference :106-137)."""
        if sinr_db < -6.0:
            return 0
        return int(min(15, np.floor((sinr_db + 6.0) / 2.0) + 1))

    def calculate_rank_indicator(self, H_channel):
        H = np.asarray(H_channel)
        ev = np.sort(np.linalg.eigvalsh(H.conj().T @ H))[::-1]
        if len(ev) >= 2:
            return 2 if ev[1] / ev[0] > 0.2 else 1
        return 1

    def generate_feedback(self, H_channel, noise_variance=1.0):
        pmi = self.calculate_pmi(H_channel)
        cqi, sinr_db = self.calculate_cqi(H_channel, pmi, noise_variance)
        return {'pmi': pmi, 'cqi': cqi, 'ri': self.calculate_rank_indicator(H_channel), 'sinr_db': sinr_db,
                'precoder': self.codebook.get_precoder(pmi)}

    def get_statistics(self):
        if not self.pmi_history:
            return None
        return {'total_feedbacks': self.total_feedbacks, 'unique_pmis': len(set(self.pmi_history)),
                'most_common_pmi': max(set(self.pmi_history), key=self.pmi_history.count),
                'pmi_distribution': np.bincount(self.pmi_history, minlength=self.codebook.codebook_size)}

    def print_statistics(self):
        stats = self.get_statistics()
        if stats is None:
            print("[CSIFeedback] No hay estadísticas disponibles")
            return
        print(f"  Total feedbacks: {stats['total_feedbacks']}")
        print(f"  PMIs únicos usados: {stats['unique_pmis']} / {self.codebook.codebook_size}")
        print(f"  PMI más común: {stats['most_common_pmi']}")
