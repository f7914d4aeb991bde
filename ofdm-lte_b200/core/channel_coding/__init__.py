"""LTE channel coding (reference core/channel_coding/__init__.py) on the kernels of csrc/coding.cu."""
from .crc import calculate_crc24a, calculate_crc24b, attach_crc24a, attach_crc24b, check_crc24a, check_crc24b
from .segmentation import segment_code_blocks, desegment_code_blocks, get_segmentation_info
from .turbo_encoder import turbo_encode, qpp_interleave, qpp_deinterleave
from .turbo_decoder import turbo_decode, LogMAPDecoder
from .rate_matching import rate_match_turbo, rate_dematching_turbo, sub_block_interleaver, sub_block_deinterleaver

__all__ = ['calculate_crc24a', 'calculate_crc24b', 'attach_crc24a', 'attach_crc24b', 'check_crc24a', 'check_crc24b',
           'segment_code_blocks', 'desegment_code_blocks', 'get_segmentation_info', 'turbo_encode', 'turbo_decode',
           'LogMAPDecoder', 'qpp_interleave', 'qpp_deinterleave', 'rate_match_turbo', 'rate_dematching_turbo',
           'sub_block_interleaver', 'sub_block_deinterleaver']
