"""autograd.rs_b200: B200-native (sm_100a) execution path for the autograd.rs hot path.

`_capi`   -- ctypes binding of the kernel-level C ABI (include/agrs.h, libagrs_b200.so)
`engine`  -- Python handles over the C++ host engine (include/agrs_engine.h, libagrs_engine.so)
             mirroring the reference's Tensor / Operator / Layer / NN / SGD API.
No CPU fallback exists anywhere in this package.
"""
from . import _capi  # noqa: F401
