"""Rank indicator / precoder selection (reference core/rank_adaptation.py:19-272).  Host-side, once
per call, on a num_rx x num_tx matrix: RI from the eigenvalues of H^H H with SNR gating, PMI by the
capacity metric over the codebook."""
import numpy as np

from .codebook_lte import LTECodebook


class RankAdaptation:
    def __init__(self, num_tx, num_rx, snr_db=15.0, rank_threshold=0.15):
        self.num_tx, self.num_rx = num_tx, num_rx
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10)
        self.rank_threshold = rank_threshold
        self.max_rank = min(num_tx, num_rx, 4)

    @staticmethod
    def _avg(H):
        return np.mean(H, axis=2) if H.ndim == 3 else H

    def calculate_optimal_rank(self, H_channel, method='eigenvalue'):
        H = self._avg(H_channel)
        if method == 'eigenvalue':
            return self._rank_from_eigenvalues(H)
        if method == 'capacity':
            return self._rank_from_capacity(H)
        raise ValueError(f"Método '{method}' no soportado")

    def _rank_from_eigenvalues(self, H):
        ev = np.sort(np.linalg.eigvalsh(H.conj().T @ H))[::-1]
        if ev[0] < 1e-10:
            return 1
        ri = min(int(np.sum(ev / ev[0] > self.rank_threshold)), self.max_rank)
        if self.snr_db < 5:
            ri = 1
        elif self.snr_db < 10:
            ri = min(ri, 2)
        return max(1, ri)

    def _rank_from_capacity(self, H):
        s = np.linalg.svd(H, compute_uv=False)[:self.max_rank]
        best_rank, best_cap = 1, -np.inf
        for rank in range(1, self.max_rank + 1):
            cap = sum(np.log2(1 + self.snr_linear * s[i] ** 2 / rank) for i in range(min(rank, len(s))))
            if cap > best_cap:
                best_cap, best_rank = cap, rank
        return best_rank

    def select_precoder_for_rank(self, H_channel, rank, metric='capacity'):
        cb = LTECodebook(self.num_tx, transmission_mode='TM4', rank=rank)
        H = self._avg(H_channel)
        best_pmi, best_val = 0, -np.inf
        for pmi in range(cb.codebook_size):
            He = H @ cb.get_precoder(pmi)
            if metric == 'capacity':
                try:
                    val = np.log2(np.linalg.det(np.eye(self.num_rx) + (self.snr_linear / rank) * (He @ He.conj().T)))
                except Exception:
                    val = 0
            elif metric == 'frobenius':
                val = np.linalg.norm(He, 'fro') ** 2
            elif metric == 'sinr':
                val = np.sum(np.abs(He) ** 2)
            else:
                raise ValueError(f"Métrica '{metric}' no soportada")
            if val > best_val:
                best_val, best_pmi = val, pmi
        return best_pmi, cb.get_precoder(best_pmi)

    def get_feedback(self, H_channel, rank_method='eigenvalue', pmi_metric='capacity'):
        ri = self.calculate_optimal_rank(H_channel, method=rank_method)
        pmi, W = self.select_precoder_for_rank(H_channel, ri, metric=pmi_metric)
        H = self._avg(H_channel)
        ev = np.sort(np.linalg.eigvalsh(H.conj().T @ H))[::-1]
        sv = np.linalg.svd(H, compute_uv=False)
        return {'ri': ri, 'pmi': pmi, 'W': W, 'eigenvalues': ev, 'condition_number': sv[0] / (sv[-1] + 1e-10)}

    def update_snr(self, new_snr_db):
        self.snr_db = new_snr_db
        self.snr_linear = 10 ** (new_snr_db / 10)
