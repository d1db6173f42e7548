"""torchlsq - B200-native drop-in for the LSQ+ fake-quantize hot path of torchlsq 2.1.

Public surface kept from the reference (/root/reference/torchlsq/__init__.py:1-17):
`LSQFakeQuantizer`, the enable_/disable_ helpers, `torchlsq.functional.lsq`,
`torchlsq.extension._HAS_OPS`.  New: `torchlsq.multi` (multi-tensor plans) and `torchlsq.dp`
(data-parallel flat gradient buffer).
"""
from .extension import _HAS_OPS

__version__ = "2.1+b200.r2"

from .quantized import *  # noqa: F401,F403,E402
