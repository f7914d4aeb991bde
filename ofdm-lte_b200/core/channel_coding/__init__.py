"""LTE channel coding (reference core/channel_coding/__init__.py) on the kernels of csrc/coding.cu: the same
public names, gathered from the stage modules."""
from . import crc, rate_matching, segmentation, turbo_decoder, turbo_encoder

_PUBLIC = {
    crc: ('calculate_crc24a', 'calculate_crc24b', 'attach_crc24a', 'attach_crc24b', 'check_crc24a', 'check_crc24b'),
    segmentation: ('segment_code_blocks', 'desegment_code_blocks', 'get_segmentation_info'),
    turbo_encoder: ('turbo_encode', 'qpp_interleave', 'qpp_deinterleave'),
    turbo_decoder: ('turbo_decode', 'LogMAPDecoder'),
    rate_matching: ('rate_match_turbo', 'rate_dematching_turbo', 'sub_block_interleaver', 'sub_block_deinterleaver'),
}
__all__ = []
for _module, _names in _PUBLIC.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)
        __all__.append(_name)
