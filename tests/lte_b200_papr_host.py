"""Loads lte_b200/papr.py's pure host helpers without importing the CUDA binding (CPU tests)."""
import importlib.util
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_ROOT, 'ofdm-lte_b200', 'lte_b200')


def _load():
    pkg = types.ModuleType('_lte_b200_host')
    pkg.__path__ = [_PKG]
    sys.modules['_lte_b200_host'] = pkg
    for name in ('sweep', 'papr'):
        spec = importlib.util.spec_from_file_location(f'_lte_b200_host.{name}', os.path.join(_PKG, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
    return sys.modules['_lte_b200_host.papr']


_papr = _load()
ccdf_from_hist = _papr.ccdf_from_hist
papr_sweep = _papr.papr_sweep
