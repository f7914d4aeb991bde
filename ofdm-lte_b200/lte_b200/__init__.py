"""lte_b200 -- B200-native engine underneath the reference-compatible API
(`config`, `core.*`, `ofdm_module` in the parent directory).

`LinkEngine` exposes the six link-chain stages as batched launches of the
hand-written sm_100a kernels in csrc/ (through the C ABI of include/lte_b200.h);
`sweep` shards Monte-Carlo trials over GPUs.
"""
from .engine import LinkEngine, chan_for  # noqa: F401
