"""ctypes binding of liblte_b200.so (include/lte_b200.h).

There is no CPU fallback: importing this module without the built library, or
calling any stage without a CUDA device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'csrc', 'liblte_b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc -gencode arch=compute_100a,code=sm_100a). lte_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

LTE_MAX_TAPS = 8
LTE_ERR_UNSUPPORTED = -2
LTE_JAKES_TONES = 16
LTE_SLOT_SYMBOLS = 14
LTE_MAX_RX = 8
LTE_MAX_TX = 8
LTE_BLK_COLS = 8
BF_MRT = 0
BF_CODEBOOK = 1
WINDOW_FULL = 0
WINDOW_USEFUL = 1


class PlanDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('Nc', C.c_int32), ('cp', C.c_int32), ('bits_per_symbol', C.c_int32),
                ('mode_simple', C.c_int32), ('num_tx_pilot_sets', C.c_int32), ('fs', C.c_double)]


class ChannelDesc(C.Structure):
    _fields_ = [('num_taps', C.c_int32), ('delay', C.c_int32 * LTE_MAX_TAPS),
                ('gain', C.c_float * LTE_MAX_TAPS), ('doppler_hz', C.c_double)]


class AwgnDesc(C.Structure):
    _fields_ = [('power', C.c_void_p), ('snr_lin', C.c_void_p), ('seed', C.c_uint64), ('row_id0', C.c_uint64),
                ('combine', C.c_int32)]


_P = C.c_void_p
_I32, _I64, _U64 = C.c_int32, C.c_int64, C.c_uint64

_SIGS = {
    'lte_version': ([], C.c_int),
    'lte_error_string': ([C.c_int], C.c_char_p),
    'lte_plan_create': ([C.POINTER(PlanDesc), _P, C.POINTER(_P)], C.c_int),
    'lte_plan_destroy': ([_P], C.c_int),
    'lte_plan_num_data': ([_P], C.c_int),
    'lte_plan_num_pilots': ([_P], C.c_int),
    'lte_plan_indices_host': ([_P, _P, _P], C.c_int),
    'lte_plan_compact_shape': ([_P, C.POINTER(_I32), C.POINTER(_I32)], C.c_int),
    'lte_plan_window': ([_P, C.c_int, C.POINTER(_I32), C.POINTER(_I32)], C.c_int),
    'lte_bits_to_indices': ([_P, _P, _I64, _P, _I64, _I32, _P], C.c_int),
    'lte_indices_to_bits': ([_P, _P, _I64, _P, _I64, _I32, _P], C.c_int),
    'lte_tx_map_ifft': ([_P, _P, _P, _I32, _P, _P, _P, _I32, _I32, _P], C.c_int),
    'lte_qam_map': ([_P, _P, _P, _I64, _P], C.c_int),
    'lte_tx_papr': ([_P, _P, _P, _I32, _P, _P, _P, _P, _P, C.c_float, C.c_float, _I32, _I32, _I32, _P], C.c_int),
    'lte_papr_symbols': ([_P, _P, _I32, _P, _P, _P, C.c_float, C.c_float, _I32, _I64, _I32, _P], C.c_int),
    'lte_histogram': ([_P, _I64, C.c_float, C.c_float, _I32, _P, _P], C.c_int),
    'lte_plan_add_dft': ([_P, _I32], C.c_int),
    'lte_dft_qam': ([_P, _P, _P, _I32, _I64, _P], C.c_int),
    'lte_equalize_zf_awgn': ([_P, _P, _P, _P, C.c_int, _I64, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_dft_m': ([_P, _P, _P, _I32, _I32, _I64, _P], C.c_int),
    'lte_channel_tdl_workspace_bytes': ([_P, C.POINTER(ChannelDesc), _I32, _I32, _I32, _I64], C.c_int64),
    'lte_channel_tdl': ([_P, C.POINTER(ChannelDesc), _P, _P, _P, _P, _P, _I32, _I32, _I32, _I64, _P], C.c_int),
    'lte_channel_rx_fft_workspace_bytes': ([_P, C.POINTER(ChannelDesc), _I32, _I32, _I32], C.c_int64),
    'lte_channel_rx_fft': ([_P, C.POINTER(ChannelDesc), _P, _P, _P, _P, _P, C.c_int, _I32, _I32, _I32, _P], C.c_int),
    'lte_channel_rx_fft_mimo_workspace_bytes': ([_P, C.POINTER(ChannelDesc), _I32, _I32, _I32, _I32], C.c_int64),
    'lte_channel_rx_fft_mimo': ([_P, C.POINTER(ChannelDesc), _P, _P, _P, _P, _P, C.c_int, _I32, _I32, _I32, _I32, _P], C.c_int),
    'lte_tx_spectral': ([_P, _P, _P, _P, _I32, _I32, _P], C.c_int),
    'lte_channel_spectral_workspace_bytes': ([_P, C.POINTER(ChannelDesc), _I32, _I32, _I32], C.c_int64),
    'lte_channel_spectral': ([_P, C.POINTER(ChannelDesc), _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P], C.c_int),
    'lte_awgn_add': ([_P, _P, _I32, _P, _P, _P, _U64, _U64, _P, _I64, _I64, _P], C.c_int),
    'lte_rx_fft': ([_P, _P, _I32, _P, _P, _P, _I32, _U64, _U64, _P, C.c_int, _I64, _I32, _P], C.c_int),
    'lte_crs_ls_interp': ([_P, _P, _P, C.c_int, C.c_int, _I64, _I32, _P], C.c_int),
    'lte_crs_ls_interp_awgn': ([_P, _P, _P, C.c_int, C.c_int, _I64, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_mrc_demap_count_awgn': ([_P, _P, _P, _P, _P, C.c_int, _I64, _I64, _I32, _I32, C.POINTER(AwgnDesc), _P],
                                 C.c_int),
    'lte_crs_ls_compact': ([_P, _P, _P, _I64, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_mrc_demap_count_compact': ([_P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_crs_mrc_demap_count_compact': ([_P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_equalize_zf': ([_P, _P, _P, _P, C.c_int, _I64, _I32, _P], C.c_int),
    'lte_equalize_mrc': ([_P, _P, _P, _P, C.c_int, _I64, _I32, _I32, _P], C.c_int),
    'lte_sfbc_encode': ([_P, _P, _P, _P, _P, _I64, _I32, _P], C.c_int),
    'lte_sfbc_decode': ([_P, _P, _P, _P, _P, C.c_int, _I64, _I32, _I32, _P], C.c_int),
    'lte_tx_sfbc_ifft': ([_P, _P, _P, _I32, _I32, _P], C.c_int),
    'lte_sfbc_decode_count': ([_P, _P, _P, _P, _P, _P, _I64, C.c_int, _I64, _I32, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_sm_precode': ([_P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _I32, _P], C.c_int),
    'lte_flat_mimo': ([_P, _P, _P, _P, _P, _I64, _I32, _I32, _I64, _P], C.c_int),
    'lte_crs_ls_pilots': ([_P, _P, _P, C.c_int, _I64, _I32, C.POINTER(AwgnDesc), _P], C.c_int),
    'lte_mimo_detect': ([_P, _P, _P, _P, _P, _I32, _I32, C.c_double, _P, _I32, _P, C.c_int, _I64, _I32, _I32, C.POINTER(AwgnDesc), _P],
                        C.c_int),
    'lte_demap_count': ([_P, _P, _P, _P, _P, _I64, _I64, _I64, _P], C.c_int),
    'lte_mrc_demap_count': ([_P, _P, _P, _P, _P, C.c_int, _I64, _I64, _I32, _I32, _P], C.c_int),
    'lte_fp32_peak_launch': ([_P, _I32, _P], C.c_int64),
    'lte_random_indices': ([_P, _P, _I64, _I64, _U64, _U64, _P], C.c_int),
    'lte_random_phases': ([_P, _I64, _I64, _U64, _U64, _P], C.c_int),
    'lte_tb_encode': ([_P, _I64, _P, _I32, _I64, _I64, _P, _P, _P, _P, _P, _P, _I64, _P], C.c_int),
    'lte_symbol_interleave': ([_P, _P, _I64, _I32, _P, _I64, _P], C.c_int),
    'lte_soft_demap': ([_P, _P, _P, C.c_int, _P, _I32, _I64, _I32, _P, _I64, _P], C.c_int),
    'lte_tb_decode_work_floats': ([_I32], C.c_int64),
    'lte_tb_decode': ([_P, _P, _I32, _I64, _I64, _I32, _P, _P, _I32, _I32, _P, _P, _P, _I64, _P, _P, _P, _P, _I64, _P],
                      C.c_int),
    'lte_crc_bits': ([_P, _I64, C.c_uint32, _I32, _P, _I64, _P], C.c_int),
    'lte_turbo_encode_blocks': ([_P, _P, _I32, _I64, _I64, _P, _P, _I64, _P], C.c_int),
    'lte_turbo_decode_blocks': ([_P, _P, _I32, _I64, _I64, _I32, _P, _I32, _I32, _P, _P, _P, _P, _I64, _P], C.c_int),
    'lte_gather_u8': ([_P, _I64, _P, _I64, _P, _I64, _P], C.c_int),
    'lte_gather_f32': ([_P, _I64, _P, _I64, _P, _I64, _P], C.c_int),
    'lte_random_channel': ([_P, _I64, _I32, _I32, _U64, _U64, _P], C.c_int),
    'lte_bf_weights': ([_P, _P, _I32, _I32, _P, _P, _P, _P, _I64, _I32, _I32, _P], C.c_int),
    'lte_rank_feedback': ([_P, _P, _P, _I32, _P, C.c_double, _I32, _P, _P, _I64, _I32, _I32, _P], C.c_int),
    'lte_bf_link': ([_P, _P, _P, _P, _P, _P, _P, _U64, _U64, _P, _P, _I64, _I64, _I32, _I32, _I32, _P], C.c_int),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)      # AttributeError here = header / library mismatch
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTS = tuple(_SIGS)


class LteError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib.lte_error_string(rc).decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise LteError(f"{what}: {msg} (code {rc})")
