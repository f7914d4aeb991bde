"""Precoder codebooks for 2 / 4 / 8 TX antennas, ranks 1-4 (reference core/codebook_lte.py:14-433;
these are the reference's simplified tables, not the TS 36.211 Householder set).  Host-side tables:
the selected W is handed to the precoding / detection kernels."""
import numpy as np


def _ph(theta):
    return np.exp(1j * theta)


class LTECodebook:
    def __init__(self, num_tx, transmission_mode='TM6', rank=1):
        self.num_tx = num_tx
        self.transmission_mode = transmission_mode
        self.rank = rank
        if transmission_mode == 'TM6' and rank != 1:
            raise ValueError(f"TM6 solo soporta rank=1, recibido rank={rank}")
        if transmission_mode == 'TM4' and (rank < 1 or rank > min(num_tx, 4)):
            raise ValueError(f"TM4 con {num_tx} antenas soporta rank 1-{min(num_tx, 4)}, recibido rank={rank}")
        if transmission_mode not in ('TM4', 'TM6'):
            raise ValueError(f"Modo {transmission_mode} no soportado")
        self.codebook = self._generate_codebook()
        self.codebook_size = len(self.codebook)

    def _generate_codebook(self):
        T, r = self.num_tx, self.rank
        cb = []
        if r == 1:                                           # reference :59-118
            if T == 2:
                cb = [np.array([[1], [v]]) / np.sqrt(2) for v in (1, -1, 1j, -1j)]
            elif T in (4, 8):
                nrm = 2 if T == 4 else np.sqrt(8)
                cb = [_ph(2 * np.pi * i * np.arange(T) / 16).reshape(-1, 1) / nrm for i in range(16)]
        elif r == 2:                                         # reference :120-207
            if T == 2:
                cb = [np.array([[1, 0], [0, 1]]), np.array([[1, 1], [1, -1]]) / np.sqrt(2),
                      np.array([[1, 1], [1j, -1j]]) / np.sqrt(2)]
            elif T == 4:
                e = [_ph(2 * np.pi * i / 4) for i in range(4)]
                cb += [np.array([[1, 0], [p, 0], [0, 1], [0, p]]) / np.sqrt(2) for p in e]
                cb += [np.array([[1, 1], [p, -p], [1, -1], [p, p]]) / 2 for p in e]
                cb += [np.array([[1, 0], [0, 1], [p, 0], [0, p]]) / np.sqrt(2) for p in e]
                cb += [np.array([[1, 1], [1, -1], [p, p], [p, -p]]) / 2 for p in e]
            elif T == 8:
                for i in range(16):
                    W = np.zeros((8, 2), dtype=complex)
                    W[0:4, 0] = W[4:8, 1] = _ph(2 * np.pi * i / 16 * np.arange(4)) / np.sqrt(4)
                    cb.append(W)
        elif r == 3:                                         # reference :209-252
            if T == 4:
                cb = [np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [p, p, p]]) / np.sqrt(2)
                      for p in (_ph(2 * np.pi * i / 8) for i in range(8))]
            elif T == 8:
                for i in range(16):
                    th = 2 * np.pi * i / 16
                    v = np.array([1, _ph(th), _ph(2 * th)]) / np.sqrt(3)
                    W = np.zeros((8, 3), dtype=complex)
                    W[0:3, 0], W[3:6, 1] = v, v
                    W[5:8, 2] = v
                    cb.append(W)
        elif r == 4:                                         # reference :254-311
            if T == 4:
                ij = np.outer(np.arange(4), np.arange(4))
                cb = [np.eye(4, dtype=complex), np.exp(-2j * np.pi * ij / 4) / 2,
                      np.array([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]) / 2,
                      np.array([[1, 1, 1, 1], [1, 1j, -1, -1j], [1, -1, 1, -1], [1, -1j, -1, 1j]]) / 2]
            elif T == 8:
                for i in range(8):
                    th = 2 * np.pi * i / 8
                    W = np.zeros((8, 4), dtype=complex)
                    for l in range(4):
                        W[2 * l:2 * l + 2, l] = np.array([1, _ph(th * (l + 1))]) / np.sqrt(2)
                    cb.append(W)
        if not cb:
            raise ValueError(f"num_tx={T} no soportado en {self.transmission_mode} Rank-{r}")
        return cb

    def get_codebook(self):
        return self.codebook

    def get_precoder(self, pmi):
        if pmi < 0 or pmi >= self.codebook_size:
            raise ValueError(f"PMI {pmi} fuera de rango [0, {self.codebook_size-1}]")
        return self.codebook[pmi]

    def select_best_pmi(self, H_channel, metric='capacity'):
        best_pmi, best_metric = 0, -np.inf
        for pmi, W in enumerate(self.codebook):
            H_eff = H_channel @ W
            if metric in ('capacity', 'sinr'):
                cur = np.sum(np.abs(H_eff) ** 2)
            elif metric == 'frobenius':
                cur = np.linalg.norm(H_eff, 'fro')
            else:
                raise ValueError(f"Métrica '{metric}' no soportada")
            if cur > best_metric:
                best_metric, best_pmi = cur, pmi
        return best_pmi, best_metric

    def calculate_quantization_error(self, H_channel, pmi):
        h_avg = np.mean(H_channel, axis=0)
        W_opt = (np.conj(h_avg) / np.linalg.norm(h_avg)).reshape(-1, 1)
        return 1 - np.abs(np.vdot(W_opt.flatten(), self.get_precoder(pmi).flatten())) ** 2

    def get_codebook_info(self):
        return {'num_tx': self.num_tx, 'transmission_mode': self.transmission_mode,
                'codebook_size': self.codebook_size, 'num_layers': self.codebook[0].shape[1],
                'pmi_bits': int(np.ceil(np.log2(self.codebook_size)))}
