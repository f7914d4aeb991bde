"""The PRODUCT numerology (ofdm-lte_b200/config.py:LTEConfig, the class users construct) against the
reference's own tables (tests/golden/tables.npz from config.py:63-154 of the unmodified reference), including the
7.5 kHz and extended-CP rows -- not only the oracle's restatement of it.  CPU only: config.py has no native part."""
import numpy as np

from helpers import load_golden


def _row(c):
    return [c.N, c.Nc, c.cp_length, c.fs]


def test_product_config_matches_reference_tables():
    from config import LTEConfig
    g = load_golden('tables')
    for bw in (1.25, 2.5, 5.0, 10.0, 15.0, 20.0):
        for cp in ('normal', 'extended'):
            assert _row(LTEConfig(bw, 15.0, 'QPSK', cp)) == list(g[f'num_{bw}_{cp}'])
    assert _row(LTEConfig(3.0, 15.0, 'QPSK')) == list(g['num_3.0_normal'])            # non-profile bandwidth
    assert _row(LTEConfig(5.0, 7.5, 'QPSK', 'extended')) == list(g['num_5.0_7.5_extended'])
    for mod, b in (('QPSK', 2), ('16-QAM', 4), ('64-QAM', 6)):
        assert LTEConfig(5.0, 15.0, mod).bits_per_symbol == b


def test_product_grid_tables_match_reference():
    """lte_b200.tables (what LinkEngine uploads into a plan) vs the reference's ResourceMapper index sets and its
    cell-id seeded pilot signs (core/resource_mapper.py:45-74, :137-152)."""
    import importlib.util
    import os
    import sys
    import types
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ofdm-lte_b200', 'lte_b200')
    # tables.py imports the package's native binding only for the ChannelDesc struct; stub it so the table code
    # can be checked on a box without the library loaded
    pkg = types.ModuleType('lte_b200_tables_pkg')
    pkg.__path__ = [root]
    sys.modules['lte_b200_tables_pkg'] = pkg
    nat = types.ModuleType('lte_b200_tables_pkg._native')
    sys.modules['lte_b200_tables_pkg._native'] = nat
    spec = importlib.util.spec_from_file_location('lte_b200_tables_pkg.tables', os.path.join(root, 'tables.py'))
    tables = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tables)
    from config import LTEConfig
    g = load_golden('tables')
    for bw in (1.25, 2.5, 5.0, 10.0, 15.0, 20.0):
        c = LTEConfig(bw, 15.0, 'QPSK')
        d, p = tables.grid_indices(c.N, c.Nc)
        assert np.array_equal(d, g[f'data_idx_{bw}']) and np.array_equal(p, g[f'pilot_idx_{bw}'])
    for cell in range(4):
        assert np.array_equal(tables.pilot_values(cell, 200), g[f'pilots_cell{cell}'])
