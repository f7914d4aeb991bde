"""Glue between the reference-shaped classes and `lte_b200.LinkEngine`.

* engines are cached per (numerology, modulation, mode, pilot layout, device);
* random draws come either from NumPy's legacy *global* RNG in the reference's exact
  order and with its re-seeding side effects (`rng='numpy'`, the default: results are
  then identical to the reference for the same global RNG state), or from the engine's
  counter-based Philox streams (`rng='philox'`, independent trials keyed by `seed`).
"""
import itertools
import threading

import numpy as np
import torch

from lte_b200 import LinkEngine
from lte_b200 import _native as nat
from lte_b200 import tables

_ENGINES = {}
_LOCK = threading.Lock()
_PHILOX_CALLS = itertools.count()


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("the lte_b200 CUDA engine needs a GPU; there is no CPU fallback")
    return torch.device('cuda', torch.cuda.current_device())


def engine_for(config, mode='lte', pilot_sets=None, bits_per_symbol=None):
    """Cached LinkEngine for an LTEConfig-like object."""
    bps = bits_per_symbol if bits_per_symbol is not None else config.bits_per_symbol
    key = (config.N, config.Nc, config.cp_length, bps, float(config.fs), mode,
           None if pilot_sets is None else np.asarray(pilot_sets).tobytes(), str(device()))
    with _LOCK:
        eng = _ENGINES.get(key)
        if eng is None:
            eng = LinkEngine(config.N, config.Nc, config.cp_length, bps, config.fs, mode=mode,
                             pilot_sets=pilot_sets, device=device())
            _ENGINES[key] = eng
    return eng


def as_bits_tensor(bits):
    """0/1 array-like (NumPy, list or torch) -> uint8 CUDA tensor [1, n]."""
    if isinstance(bits, torch.Tensor):
        t = bits.to(device=device(), dtype=torch.uint8)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(bits).astype(np.uint8))).to(device())
    return t.reshape(1, -1)


def as_complex_tensor(x):
    if isinstance(x, torch.Tensor):
        return x.to(device=device(), dtype=torch.complex64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex64))).to(device())


def to_numpy(t, dtype=None):
    a = t.detach().cpu().numpy()
    return a.astype(dtype) if dtype is not None else a


def reference_pilot_side_effect(cell_id, num_pilots):
    """The reference generates its pilots with np.random.seed(cell_id) followed by
    np.random.choice (core/resource_mapper.py:148-149), re-seeding the caller's global RNG
    on every map_symbols() / estimate_channel().  Kept so `rng='numpy'` stays draw-for-draw
    compatible."""
    np.random.seed(cell_id)
    np.random.choice([1, -1], size=num_pilots)


class NumpyDraws:
    """Draws from NumPy's legacy global RNG in the reference's order
    (core/rayleighchannel.py:31: rand(16) per tap; core/channel.py:58-59: normal() real part
    then imaginary part)."""
    kind = 'numpy'

    def phases(self, n_taps):
        if n_taps == 0:
            return None
        u = np.stack([np.random.rand(nat.LTE_JAKES_TONES) for _ in range(n_taps)])
        return torch.from_numpy(u.astype(np.float32)).to(device())

    def unit_normals(self, n):
        zr = np.random.normal(0.0, 1.0, n)
        zi = np.random.normal(0.0, 1.0, n)
        return torch.from_numpy((zr + 1j * zi).astype(np.complex64)).to(device())


class PhiloxDraws:
    """Engine-side counter-based draws: every call of a simulate_* method gets a fresh
    stream id, so repeated calls are independent trials."""
    kind = 'philox'

    def __init__(self, seed=0):
        self.seed = int(seed)

    def next_stream(self):
        return next(_PHILOX_CALLS)


def make_draws(rng, seed):
    if rng == 'numpy':
        return NumpyDraws()
    if rng == 'philox':
        return PhiloxDraws(seed)
    raise ValueError(f"rng must be 'numpy' or 'philox', got {rng!r}")
