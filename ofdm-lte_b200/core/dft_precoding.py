"""SC-FDM DFT precoding / IDFT decoding (reference core/dft_precoding.py:20-348).

The reference multiplies by dense M x M matrices; here every call is one launch of the
Bluestein M-point DFT kernel (csrc/dft.cu, `lte_dft_m`)."""
from typing import Dict

import numpy as np
import torch

from . import _backend as be


def _dft(symbols, M, inverse):
    from config import LTEConfig
    eng = be.engine_for(LTEConfig())          # the DFT kernel only needs a plan handle
    x = be.as_complex_tensor(symbols).reshape(-1)
    y = eng.dft_m(x, M, inverse=inverse)
    return y if isinstance(symbols, torch.Tensor) else be.to_numpy(y)


class DFTPrecodifier:
    def __init__(self, M: int = None, enable: bool = True):
        self.M, self.enable = M, enable

    def set_size(self, M: int):
        self.M = M

    def precoding(self, symbols):
        if not self.enable or self.M is None:
            return symbols
        if len(symbols) != self.M:
            raise ValueError(f"Tamaño de símbolos ({len(symbols)}) debe ser igual a M ({self.M})")
        return _dft(symbols, self.M, False)

    precoding_ifft = precoding

    def get_statistics(self) -> Dict:
        return {'enabled': self.enable, 'dft_size': self.M, 'matrix_computed': False}


class IDFTDecodifier:
    def __init__(self, M: int = None, enable: bool = True):
        self.M, self.enable = M, enable

    def set_size(self, M: int):
        self.M = M

    def decoding(self, precoded_symbols):
        if not self.enable or self.M is None:
            return precoded_symbols
        if len(precoded_symbols) != self.M:
            raise ValueError(f"Tamaño de símbolos ({len(precoded_symbols)}) debe ser igual a M ({self.M})")
        return _dft(precoded_symbols, self.M, True)

    decoding_fft = decoding

    def get_statistics(self) -> Dict:
        return {'enabled': self.enable, 'idft_size': self.M, 'matrix_computed': False}


class SC_FDMPrecodifier:
    """reference core/dft_precoding.py:254-300."""

    def __init__(self, num_data_subcarriers: int, enable: bool = True):
        self.num_data_subcarriers = num_data_subcarriers
        self.enable = enable
        self.dft_precoder = DFTPrecodifier(M=num_data_subcarriers, enable=enable)

    def precoding(self, data_symbols):
        if not self.enable:
            return data_symbols
        return self.dft_precoder.precoding(data_symbols)

    def set_enable(self, enable: bool):
        self.enable = enable
        self.dft_precoder.enable = enable

    def get_statistics(self) -> Dict:
        return {'sc_fdm_enabled': self.enable, 'num_data_subcarriers': self.num_data_subcarriers,
                'dft_stats': self.dft_precoder.get_statistics()}


class SC_FDMDecodifier:
    """reference core/dft_precoding.py:303-348."""

    def __init__(self, num_data_subcarriers: int, enable: bool = True):
        self.num_data_subcarriers = num_data_subcarriers
        self.enable = enable
        self.idft_decoder = IDFTDecodifier(M=num_data_subcarriers, enable=enable)

    def decoding(self, precoded_symbols):
        if not self.enable:
            return precoded_symbols
        return self.idft_decoder.decoding(precoded_symbols)

    def set_enable(self, enable: bool):
        self.enable = enable
        self.idft_decoder.enable = enable

    def get_statistics(self) -> Dict:
        return {'sc_fdm_enabled': self.enable, 'num_data_subcarriers': self.num_data_subcarriers,
                'idft_stats': self.idft_decoder.get_statistics()}
