"""gmmvi_b200 -- B200-native backend for the GMMVI (SAMTRON) hot path.

Mirrors the module API of OlegArenz/gmmvi (models, optimization, configs, gmmvi_runner) on torch CUDA
tensors; all numerical work runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/gmmvi_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"


def _configure_allocator():
    """Device-memory policy.  SAMTRON's adaptive number of components makes every [K, N]-shaped temporary grow a little
    each iteration (examples/6 adds a component per iteration); with torch's default caching allocator a request that
    is larger than every cached block turns into a synchronising cudaMalloc -- measured 5 ms of host time per C2
    iteration (profiles/r02_host_profile_c2_before.txt, 923 x torch.empty = 0.17 s of 0.32 s).  Expandable segments
    grow the existing mapping instead.  A PYTORCH_CUDA_ALLOC_CONF set by the user wins."""
    import os
    if "PYTORCH_CUDA_ALLOC_CONF" in os.environ or "PYTORCH_ALLOC_CONF" in os.environ:
        return
    try:
        import torch
        setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or torch.cuda.memory._set_allocator_settings
        setter("expandable_segments:True")
    except Exception:      # older torch / no CUDA build: keep the default policy
        pass


_configure_allocator()
