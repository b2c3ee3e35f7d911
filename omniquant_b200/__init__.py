"""omniquant_b200 -- B200-native (sm_100a) implementation of OmniQuant's block-wise
calibration hot path (LWC + LET) behind the reference's module API.

Device compute goes through the C-ABI library ``lib/liboq_b200.so``
(``include/omniquant_b200.h``); PyTorch provides device memory, streams and
``torch.distributed`` only.
"""
__version__ = "0.1.0"
