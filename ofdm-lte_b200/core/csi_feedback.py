"""CSI feedback (PMI / CQI / RI) for the beamforming path -- the reference's core/csi_feedback.py:25-228.

The per-channel work (codebook search for the PMI, effective channel, post-precoding power) is the
`lte_bf_weights` kernel: one thread per channel matrix in fp64 registers, so `feedback_batch` serves a whole
Monte-Carlo batch in one launch and the scalar methods are the B = 1 case of it.  CQI quantisation and the
rank indicator are bookkeeping on the numbers that come back."""
import collections

import numpy as np

from . import _backend as be
from .codebook_lte import LTECodebook

_CQI_FLOOR_DB, _CQI_STEP_DB, _CQI_MAX = -6.0, 2.0, 15       # reference table :106-137: 2 dB bins from -6 dB
_RI_RATIO = 0.2                                             # second / first eigenvalue above which rank 2 is reported


class _BinsConfig:
    """Smallest plan that can carry the precoder kernels (they never touch the grid tables)."""
    N, Nc, cp_length, fs, bits_per_symbol = 128, 76, 9, 1.92e6, 2


def _bf_engine():
    return be.engine_for(_BinsConfig)


def _as_batch(H_channel):
    """[R, T] (or a single row) -> complex64 CUDA tensor [1, R, T]."""
    H = np.atleast_2d(np.asarray(H_channel, dtype=np.complex64))
    return be.as_complex_tensor(H[None])


class CSIFeedback:
    def __init__(self, num_tx, num_rx, codebook_type='TM6', feedback_mode='perfect'):
        self.num_tx, self.num_rx = num_tx, num_rx
        self.codebook_type, self.feedback_mode = codebook_type, feedback_mode
        self.codebook = LTECodebook(num_tx, transmission_mode=codebook_type)
        self.pmi_history = []
        self.total_feedbacks = 0

    # ---- GPU entry ---------------------------------------------------------------------------------
    def feedback_batch(self, h):
        """h: complex64 CUDA tensor [B, R, T] -> (pmi int32 [B], H_eff = H W_pmi [B, R])."""
        _, heff, pmi, _ = _bf_engine().bf_weights(h, self.codebook.codebook, mode='CODEBOOK')
        return pmi, heff

    # ---- the reference's scalar API -----------------------------------------------------------------
    def calculate_pmi(self, H_channel):
        pmi = int(self.feedback_batch(_as_batch(H_channel))[0].item())
        self.pmi_history.append(pmi)
        self.total_feedbacks += 1
        return pmi

    def calculate_cqi(self, H_channel, pmi, noise_variance=1.0):
        gain = np.linalg.norm(np.asarray(H_channel) @ self.codebook.get_precoder(pmi)) ** 2
        sinr_db = 10 * np.log10(gain / noise_variance)
        return self._sinr_to_cqi(sinr_db), sinr_db

    @staticmethod
    def _sinr_to_cqi(sinr_db):
        if sinr_db < _CQI_FLOOR_DB:
            return 0
        return int(min(_CQI_MAX, 1 + (sinr_db - _CQI_FLOOR_DB) // _CQI_STEP_DB))

    def calculate_rank_indicator(self, H_channel):
        H = np.asarray(H_channel)
        lam = np.linalg.eigvalsh(H.conj().T @ H)[::-1]            # descending
        return 2 if len(lam) > 1 and lam[1] / lam[0] > _RI_RATIO else 1

    def generate_feedback(self, H_channel, noise_variance=1.0):
        pmi = self.calculate_pmi(H_channel)
        cqi, sinr_db = self.calculate_cqi(H_channel, pmi, noise_variance)
        return dict(pmi=pmi, cqi=cqi, ri=self.calculate_rank_indicator(H_channel), sinr_db=sinr_db,
                    precoder=self.codebook.get_precoder(pmi))

    # ---- statistics ---------------------------------------------------------------------------------
    def get_statistics(self):
        if not self.pmi_history:
            return None
        counts = collections.Counter(self.pmi_history)
        top = max(sorted(counts), key=counts.get)
        return dict(total_feedbacks=self.total_feedbacks, unique_pmis=len(counts), most_common_pmi=top,
                    pmi_distribution=np.bincount(self.pmi_history, minlength=self.codebook.codebook_size))

    def print_statistics(self):
        st = self.get_statistics()
        if st is None:
            print("[CSIFeedback] No hay estadísticas disponibles")
            return
        print(f"[CSIFeedback] {st['total_feedbacks']} feedbacks, {st['unique_pmis']} / "
              f"{self.codebook.codebook_size} PMIs, most common {st['most_common_pmi']}")
