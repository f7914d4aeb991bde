"""Shared helpers for the parity tests: golden loading and oracle invocation."""
import os

import numpy as np

from oracle import lte_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel_err(a, b):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_bits(g):
    return np.unpackbits(g['bits'])[:int(g['nbits'])].astype(np.int64)


def golden_bits_rx(g, snr):
    return np.unpackbits(g[f'bits_rx_{snr}'])[:int(g['nbits'])].astype(np.int64)


def numerology(case):
    return O.Numerology(case['bw'], 15.0, case['mod'], case.get('cp_type', 'normal'))


def oracle_siso(case, bits, snr, **kw):
    if 'global_seed' in case and 'z' not in kw:
        kw['draws'] = O.ReferenceDraws(0, global_seed=case['global_seed'])
    return O.simulate_siso(bits, snr, numerology(case), channel_type=case['ch'],
                           mode=case.get('mode', 'lte'), sc_fdm=case.get('sc_fdm', False),
                           equalize=case.get('equalize', True), itu_profile=case['prof'],
                           frequency_ghz=2.0, velocity_kmh=case['v'], **kw)


def oracle_simo(case, bits, snr, **kw):
    return O.simulate_simo(bits, snr, numerology(case), num_rx=case['R'], channel_type=case['ch'],
                           itu_profile=case['prof'], frequency_ghz=2.0, velocity_kmh=case['v'], **kw)


def reference_draws(case, n_samples, n_links):
    """Phases [links][taps][16] and unit normals [links][2][n] in the reference's
    draw order (see oracle.lte_oracle.ReferenceDraws)."""
    num = numerology(case)
    _, pilot_idx = O.grid_indices(num.N, num.Nc)
    d = O.ReferenceDraws(len(pilot_idx))
    n_taps = len(O.ITU[case['prof']][0]) if case['ch'] == 'rayleigh_mp' else 0
    phases, z = [], []
    for _ in range(n_links):
        phases.append(d.phases(n_taps) if n_taps else np.zeros((0, 16)))
        z.append(np.stack(d.unit_normals(n_samples)))
    return np.stack(phases), np.stack(z)
