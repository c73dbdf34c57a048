"""TEST INFRASTRUCTURE: the CPU stand-in of the C ABI (tests/cpp/standin_dsdgpu.cpp) bound
through the product's own ctypes prototypes, so multi-rank host logic can run on a CPU box."""
import ctypes as C
import functools
import os

from dsf2flac_b200 import capi

CPU_HOST = os.path.join(os.path.dirname(__file__), "cpp", "_build", "libdsdhost_cpu.so")
_NAMES = ["dsdgpu_create", "dsdgpu_destroy", "dsdgpu_last_error", "dsdgpu_set_option", "dsdgpu_kernel_launches",
          "dsdgpu_decimate_host", "dsdgpu_run_batch", "dsdgpu_plan_shards"]


@functools.lru_cache(None)
def lib():
    return capi.bind(C.CDLL(CPU_HOST), _NAMES)


def make_ctx(lut, nch, nstep, device=0, lut_precision=64):
    return capi.DsdGpu(lut, nch, nstep, device=device, lut_precision=lut_precision, lib=lib())


def cpu_host_lib():
    """The drop-in host library linked against the CPU stand-in (dsdhost_* drivers)."""
    return capi.host_lib(CPU_HOST)
