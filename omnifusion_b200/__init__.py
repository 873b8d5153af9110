"""omnifusion_b200: B200-native (sm_100a) implementation of OmniFusion's tangent-patch
inference path behind the reference's own interfaces.

    from omnifusion_b200.equi_pers.equi2pers_v3 import equi2pers
    from omnifusion_b200.equi_pers.pers2equi_v3 import pers2equi
    from omnifusion_b200.model.spherical_model_iterative import spherical_fusion

All arithmetic runs in hand-written CUDA kernels in libofb.so (C ABI: include/ofb.h),
loaded through ctypes; PyTorch only provides device memory and streams.  There is no CPU
fallback.
"""
__version__ = "0.1.0"
