"""CRC-24A / CRC-24B / CRC-16 (reference core/channel_coding/crc.py:32-365) on `lte_crc_bits`: a shift register
per row, zero initial state, MSB first."""
import numpy as np

from lte_b200 import _native as nat
from . import _gpu as g

CRC24A_POLYNOMIAL = 0x1864CFB
CRC24B_POLYNOMIAL = 0x1800063
CRC16_POLYNOMIAL = 0x11021


def _calculate_crc(data_bits, polynomial, crc_length):
    b = g.bits_dev(np.asarray(data_bits)) if len(data_bits) else None
    import torch
    out = torch.empty((1, crc_length), dtype=torch.uint8, device=g.be.device())
    if b is None:
        return np.zeros(crc_length, dtype=np.uint8)
    nat.check(nat.lib.lte_crc_bits(g.ptr(b), b.shape[1], polynomial & ((1 << crc_length) - 1), crc_length, g.ptr(out), 1,
                                   g.stream()), 'lte_crc_bits')
    return out.cpu().numpy()[0]


def calculate_crc24a(data_bits):
    return _calculate_crc(data_bits, CRC24A_POLYNOMIAL, 24)


def calculate_crc24b(data_bits):
    return _calculate_crc(data_bits, CRC24B_POLYNOMIAL, 24)


def calculate_crc16(data_bits):
    return _calculate_crc(data_bits, CRC16_POLYNOMIAL, 16)


def attach_crc24a(data_bits):
    return np.concatenate([np.asarray(data_bits), calculate_crc24a(data_bits)])


def attach_crc24b(data_bits):
    return np.concatenate([np.asarray(data_bits), calculate_crc24b(data_bits)])


def attach_crc16(data_bits):
    return np.concatenate([np.asarray(data_bits), calculate_crc16(data_bits)])


def _check(data_with_crc, fn, n):
    d = np.asarray(data_with_crc)
    if len(d) < n:
        return False
    return bool(np.array_equal(d[-n:], fn(d[:-n])))


def check_crc24a(data_with_crc):
    return _check(data_with_crc, calculate_crc24a, 24)


def check_crc24b(data_with_crc):
    return _check(data_with_crc, calculate_crc24b, 24)


def check_crc16(data_with_crc):
    return _check(data_with_crc, calculate_crc16, 16)
