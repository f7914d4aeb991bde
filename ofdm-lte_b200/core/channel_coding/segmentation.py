"""Code-block segmentation (reference core/channel_coding/segmentation.py:28-459).  The block sizes, filler and
bit split come from lte_b200.coding.block_layout (the layout the kernels use); the bit movement of one transport
block is a gather, the per-block CRC-24B is `lte_crc_bits`."""
import numpy as np

from lte_b200.coding import block_layout, interleaver_size
from lte_b200.qpp_table import TURBO_K
from . import _gpu as g
from .crc import calculate_crc24b

TURBO_INTERLEAVER_SIZES = list(TURBO_K)


def find_interleaver_size(min_size):
    return interleaver_size(min_size)


def _meta(B, layout):
    seg = len(layout) > 1 or bool(layout[0][3])
    fill = [f for _, f, _, _ in layout]
    m = {'num_blocks': len(layout), 'block_sizes': [k for k, _, _, _ in layout], 'num_filler_bits': int(sum(fill)),
         'filler_positions': list(range(int(sum(fill)))), 'filler_per_block': fill, 'original_size': B, 'segmented': seg}
    if seg:
        kp = max(m['block_sizes'])
        i = TURBO_INTERLEAVER_SIZES.index(kp)
        m.update(K_plus=kp, K_minus=TURBO_INTERLEAVER_SIZES[i - 1] if i > 0 else kp,
                 C_plus=m['block_sizes'].count(kp), C_minus=len(layout) - m['block_sizes'].count(kp))
    return m


def segment_code_blocks(transport_block_with_crc):
    tb = np.asarray(transport_block_with_crc)
    B = len(tb)
    layout = block_layout(B)
    # one gather lays the info bits of every block behind its filler zeros (-1 = constant 0)
    table, pos = [], 0
    for K, F, n, has_crc in layout:
        body = K - (24 if has_crc else 0)
        table += [-1] * F + list(range(pos, pos + n)) + [-1] * (body - F - n)
        pos += n
    flat = g.gather_any(tb.astype(np.uint8), np.array(table, dtype=np.int32))
    blocks, off = [], 0
    for K, F, n, has_crc in layout:
        body = K - (24 if has_crc else 0)
        cb = flat[off:off + body]
        off += body
        blocks.append(np.concatenate([cb, calculate_crc24b(cb)]).astype(np.uint8) if has_crc else cb.astype(np.uint8))
    return blocks, _meta(B, layout)


def desegment_code_blocks(code_blocks, metadata):
    B = metadata['original_size']
    layout = block_layout(B)
    out = [np.asarray(cb)[F:F + n] for cb, (K, F, n, has_crc) in zip(code_blocks, layout)]
    return np.concatenate(out) if len(out) > 1 else out[0]


def get_segmentation_info(transport_block_size):
    ks = [k for k, _, _, _ in block_layout(transport_block_size)]
    return {'num_blocks': len(ks), 'block_sizes': ks, 'total_coded_bits': sum(3 * k + 12 for k in ks)}
