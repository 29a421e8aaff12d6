"""gmmvi_b200 -- B200-native backend for the GMMVI (SAMTRON) hot path.

Mirrors the module API of OlegArenz/gmmvi (models, optimization, configs, gmmvi_runner) on torch CUDA
tensors; all numerical work runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/gmmvi_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"

