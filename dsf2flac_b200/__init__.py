"""dsf2flac_b200 -- the B200 (sm_100a) DSD->PCM decimation path of dsf2flac.

Product:  libdsdgpu.so  (csrc/: CUDA kernels + the C ABI of include/dsdgpu.h)
          libdsdhost.so (host/: link-time drop-in for the reference's DsdDecimator class)
Python here is plumbing for tests and benchmarks: capi (ctypes binding), filters (the
reference's filter geometry and table builder, host side), batch (one-process-per-GPU driver).
"""
from . import capi, filters  # noqa: F401
