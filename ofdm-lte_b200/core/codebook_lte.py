"""Precoder codebooks for 2 / 4 / 8 TX antennas, ranks 1-4 (reference core/codebook_lte.py:14-433;
these are the reference's simplified tables, not the TS 36.211 Householder set).  Host-side tables:
the selected W is handed to the precoding / detection kernels."""
import numpy as np


def _ph(theta):
    return np.exp(1j * theta)


def _cols(*columns):
    """Column vectors -> [T, L] matrix."""
    return np.stack([np.asarray(c, dtype=complex) for c in columns], axis=1)


def _steer(n, step, count):
    """DFT steering vector: exp(j step m), m = 0..n-1, for `count` equally spaced steps of 2 pi / count."""
    return [_ph(2 * np.pi * i / count * np.arange(n)) for i in range(count)]


# ---- rank 1 (TM6 vectors, re-used by TM4 rank 1; reference :59-118) ---------------------------------------
def _r1_t2():
    return [_cols([1, v]) / np.sqrt(2) for v in (1, -1, 1j, -1j)]


def _r1_dft(T):
    return lambda: [_cols(v) / np.sqrt(T) for v in _steer(T, None, 16)]


# ---- rank 2 (reference :120-207) ---------------------------------------------------------------------------
def _r2_t2():
    return [np.eye(2), _cols([1, 1], [1, -1]) / np.sqrt(2), _cols([1, 1j], [1, -1j]) / np.sqrt(2)]


def _r2_t4():
    ph = [_ph(2 * np.pi * i / 4) for i in range(4)]
    groups = ([_cols([1, p, 0, 0], [0, 0, 1, p]) / np.sqrt(2) for p in ph],
              [_cols([1, p, 1, p], [1, -p, -1, p]) / 2 for p in ph],
              [_cols([1, 0, p, 0], [0, 1, 0, p]) / np.sqrt(2) for p in ph],
              [_cols([1, 1, p, p], [1, -1, p, -p]) / 2 for p in ph])
    return [W for g in groups for W in g]


def _r2_t8():
    out = []
    for v in _steer(4, None, 16):
        W = np.zeros((8, 2), dtype=complex)
        W[:4, 0] = W[4:, 1] = v / 2
        out.append(W)
    return out


# ---- rank 3 (reference :209-252) ---------------------------------------------------------------------------
def _r3_t4():
    return [_cols([1, 0, 0, p], [0, 1, 0, p], [0, 0, 1, p]) / np.sqrt(2) for p in (_ph(2 * np.pi * i / 8) for i in range(8))]


def _r3_t8():
    out = []
    for v in _steer(3, None, 16):
        W = np.zeros((8, 3), dtype=complex)
        for col, row in enumerate((0, 3, 5)):
            W[row:row + 3, col] = v / np.sqrt(3)
        out.append(W)
    return out


# ---- rank 4 (reference :254-311) ---------------------------------------------------------------------------
def _r4_t4():
    m = np.arange(4)
    had = np.array([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]])
    quad = np.array([[1, 1, 1, 1], [1, 1j, -1, -1j], [1, -1, 1, -1], [1, -1j, -1, 1j]])
    return [np.eye(4), np.exp(-2j * np.pi * np.outer(m, m) / 4) / 2, had / 2, quad / 2]


def _r4_t8():
    out = []
    for i in range(8):
        W = np.zeros((8, 4), dtype=complex)
        for layer in range(4):
            W[2 * layer:2 * layer + 2, layer] = np.array([1, _ph(2 * np.pi * i / 8 * (layer + 1))]) / np.sqrt(2)
        out.append(W)
    return out


_BUILDERS = {(1, 2): _r1_t2, (1, 4): _r1_dft(4), (1, 8): _r1_dft(8), (2, 2): _r2_t2, (2, 4): _r2_t4, (2, 8): _r2_t8,
             (3, 4): _r3_t4, (3, 8): _r3_t8, (4, 4): _r4_t4, (4, 8): _r4_t8}


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
        """Tables keyed by (rank, num_tx); each builder returns the list of [T, rank] precoders."""
        build = _BUILDERS.get((self.rank, self.num_tx))
        if build is None:
            raise ValueError(f"num_tx={self.num_tx} no soportado en {self.transmission_mode} Rank-{self.rank}")
        return [np.asarray(W, dtype=complex) for W in build()]

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
