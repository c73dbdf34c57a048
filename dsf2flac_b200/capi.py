"""ctypes binding of the C ABI in include/dsdgpu.h (dsf2flac_b200/libdsdgpu.so) and of the
C-callable drivers around the drop-in DsdDecimator (dsf2flac_b200/libdsdhost.so).

This is the only way Python reaches the product: tests/, bench.py and __graft_entry__.py all
go through these calls.  There is no Python or CPU implementation of the operator behind them;
if the CUDA library is missing, loading fails loudly.

The operator is the body of the reference's DsdDecimator::getSamplesInternal
(reference src/dsd_decimator.cpp:173-222); see include/dsdgpu.h for the exact definition.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(_HERE, "libdsdgpu.so")
HOST_LIB = os.path.join(_HERE, "libdsdhost.so")

OUT_INT32, OUT_FLOAT64, OUT_PCM24, OUT_PCM16 = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_RING, KERNEL_SEGMENTED = 0, 1, 3, 4
ERR_NODEVICE = 3

_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in step with include/dsdgpu.h (tests/test_abi.py checks that
# every prototype in the header is bound here and exported by the library)
GPU_PROTOTYPES = {
    "dsdgpu_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "dsdgpu_destroy": (None, [_vp]),
    "dsdgpu_last_error": (C.c_char_p, [_vp]),
    "dsdgpu_set_option": (C.c_int, [_vp, C.c_char_p, C.c_longlong]),
    "dsdgpu_kernel_launches": (C.c_ulonglong, [_vp]),
    "dsdgpu_host_alloc": (_vp, [C.c_size_t]),
    "dsdgpu_host_free": (None, [_vp]),
    "dsdgpu_host_register": (C.c_int, [_vp, C.c_size_t]),
    "dsdgpu_host_unregister": (C.c_int, [_vp]),
    "dsdgpu_decimate_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64,
                                         C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _vp, _vp]),
    "dsdgpu_decimate_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64,
                                       C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _vp]),
    "dsdgpu_submit": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64,
                                C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _vp]),
    "dsdgpu_wait": (C.c_int, [_vp, C.c_int]),
    "dsdgpu_dither_fill_device": (C.c_int, [_vp, C.c_uint64, C.c_int64, _vp, _vp]),
    "dsdgpu_dop_pack_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64, C.c_int,
                                         _vp, _vp]),
    "dsdgpu_dop_pack_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64, C.c_int,
                                       _vp]),
    "dsdgpu_run_batch": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_char_p, C.c_size_t]),
    "dsdgpu_plan_shards": (C.c_int64, [C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                       C.c_int, C.c_int, C.c_int64, _vp, C.c_int64]),
}

_gpu = None
_host = None


def gpu_lib():
    """libdsdgpu.so, loaded once.  Raises if it has not been built (make gpu)."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB):
            raise RuntimeError(f"{GPU_LIB} is missing: build it with `make gpu` (there is no fallback path)")
        _gpu = bind(C.CDLL(GPU_LIB))
    return _gpu


def bind(lib, names=None):
    """Attach the prototypes of include/dsdgpu.h to a loaded library (tests bind a CPU stand-in
    of the ABI this way to exercise host logic; the product only ever binds libdsdgpu.so)."""
    for name in (names or GPU_PROTOTYPES):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = GPU_PROTOTYPES[name]
    return lib


def host_lib(path=None):
    """libdsdhost.so (drop-in DsdDecimator + drivers); `path` lets a test load another build."""
    global _host
    if path is None and _host is not None:
        return _host
    p = path or HOST_LIB
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `make host`")
    lib = C.CDLL(p)
    i64, u32, dbl = C.c_int64, C.c_uint32, C.c_double
    lib.dsdhost_last_error.restype = C.c_char_p
    lib.dsdhost_last_error.argtypes = []
    lib.dsdhost_set_extension_filters.restype = None
    lib.dsdhost_set_extension_filters.argtypes = [C.c_int]
    lib.dsdhost_info.restype = C.c_int
    lib.dsdhost_info.argtypes = [u32, u32, i64, C.POINTER(C.c_int32), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)]
    lib.dsdhost_run_raw.restype = C.c_int
    lib.dsdhost_run_raw.argtypes = [_vp, C.c_int, i64, i64, C.c_uint8, C.c_int, u32, u32, i64, i64, i64, dbl, dbl, dbl,
                                    C.c_int, u32, _vp, _vp]
    lib.dsdhost_run_raw_typed.restype = C.c_int
    lib.dsdhost_run_raw_typed.argtypes = [_vp, C.c_int, i64, i64, C.c_uint8, C.c_int, u32, u32, i64, i64, i64, dbl, dbl, dbl,
                                          C.c_int, u32, C.c_int, _vp]
    lib.dsdhost_run_app.restype = i64
    lib.dsdhost_run_app.argtypes = [_vp, C.c_int, i64, i64, C.c_uint8, C.c_int, u32, u32, dbl, dbl, dbl, C.c_int, u32,
                                    _vp, i64]
    lib.dsdhost_run_app_bulk.restype = i64
    lib.dsdhost_run_app_bulk.argtypes = lib.dsdhost_run_app.argtypes
    lib.dsdhost_run_script.restype = i64
    lib.dsdhost_run_script.argtypes = [_vp, C.c_int, i64, i64, C.c_uint8, C.c_int, u32, u32, C.c_int, C.POINTER(i64), C.c_int,
                                       dbl, dbl, dbl, dbl, _vp, i64]
    lib.dsdhost_run_precision_after_bulk.restype = i64
    lib.dsdhost_run_precision_after_bulk.argtypes = [_vp, C.c_int, i64, i64, i64, C.c_int, u32, u32, C.c_int, dbl, dbl, dbl, i64, _vp]
    lib.dsdhost_file_info.restype = C.c_int
    lib.dsdhost_file_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(i64), C.POINTER(i64)]
    lib.dsdhost_file_channel_bytes.restype = i64
    lib.dsdhost_file_channel_bytes.argtypes = [C.c_char_p, C.c_int, C.c_int, _vp, i64]
    lib.dsdhost_convert_file.restype = i64
    lib.dsdhost_convert_file.argtypes = [C.c_char_p, C.c_int, u32, dbl, dbl, dbl, C.c_int, u32, _vp, i64]
    lib.dsdhost_dop_run_app.restype = i64
    lib.dsdhost_dop_run_app.argtypes = [_vp, C.c_int, i64, i64, C.c_uint8, C.c_int, _vp, i64]
    lib.dsdhost_dop_file.restype = i64
    lib.dsdhost_dop_file.argtypes = [C.c_char_p, C.c_int, _vp, i64]
    lib.dsdhost_synth_sdm.restype = None
    lib.dsdhost_synth_sdm.argtypes = [_vp, i64, dbl, dbl, dbl, C.c_int]
    if path is None:
        _host = lib
    return lib


class DsdGpuError(RuntimeError):
    pass


def _ptr(x):
    """Raw address of a numpy array, a torch tensor or an int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()          # torch.Tensor


class DsdGpu:
    """One dsdgpu_ctx: the tables of one decimator on one GPU.

    lut: float64 array [nlut, 256] built exactly as the reference's initLookupTable does
    (dsd_decimator.cpp:117-151); nstep = ratio/8.
    """

    def __init__(self, lut, nch, nstep, device=0, lut_precision=64, lib=None):
        self.lib = lib or gpu_lib()
        lut = np.ascontiguousarray(lut, dtype=np.float64)
        assert lut.ndim == 2 and lut.shape[1] == 256
        self.nch, self.nlut, self.nstep = int(nch), int(lut.shape[0]), int(nstep)
        h = _vp()
        rc = self.lib.dsdgpu_create(C.byref(h), device, self.nch, self.nlut, self.nstep,
                                    lut.ctypes.data_as(C.POINTER(C.c_double)), lut_precision)
        if rc != 0:
            raise DsdGpuError(f"dsdgpu_create failed ({rc}): {self.lib.dsdgpu_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsdgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise DsdGpuError(f"{what} failed ({rc}): {self.lib.dsdgpu_last_error(self.h).decode()}")

    def set_option(self, key, value):
        self._check(self.lib.dsdgpu_set_option(self.h, key.encode(), int(value)), f"set_option({key})")

    @property
    def launches(self):
        return int(self.lib.dsdgpu_kernel_launches(self.h))

    def decimate_device(self, d_in, in_stride, in_first, in_count, fill, p0, nframes, scale, dither_amp, clip,
                        dither_first_sample, out_kind, d_out, stream=None):
        """Device pointers in and out (torch tensors or raw addresses); asynchronous on `stream`
        (a raw cudaStream_t value) or synchronous on the context's stream when stream is None."""
        self._check(self.lib.dsdgpu_decimate_device(self.h, _ptr(d_in), in_stride, in_first, in_count, fill, p0,
                                                    nframes, scale, dither_amp, clip, dither_first_sample, out_kind,
                                                    _ptr(d_out), stream), "dsdgpu_decimate_device")

    def decimate_host(self, planar, p0, nframes, scale, dither_amp=0.0, clip=0.0, dither_first_sample=0,
                      out_kind=OUT_INT32, fill=0x69, in_first=0, out=None):
        """planar: uint8 [nch, nbytes] host array holding stream bytes in_first.. of every channel."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        assert planar.ndim == 2 and planar.shape[0] == self.nch
        if out is None:
            out = (np.empty((nframes, self.nch, 3), dtype=np.uint8) if out_kind == OUT_PCM24 else
                   np.empty((nframes, self.nch), dtype={OUT_INT32: np.int32, OUT_FLOAT64: np.float64, OUT_PCM16: np.int16}[out_kind]))
        self._check(self.lib.dsdgpu_decimate_host(self.h, planar.ctypes.data, planar.shape[1], in_first,
                                                  planar.shape[1], fill, p0, nframes, scale, dither_amp, clip,
                                                  dither_first_sample, out_kind, out.ctypes.data),
                    "dsdgpu_decimate_host")
        return out

    def decimate_host_raw(self, in_ptr, in_stride, in_first, in_count, fill, p0, nframes, scale, dither_amp, clip,
                          dither_first_sample, out_kind, out_ptr):
        self._check(self.lib.dsdgpu_decimate_host(self.h, in_ptr, in_stride, in_first, in_count, fill, p0, nframes,
                                                  scale, dither_amp, clip, dither_first_sample, out_kind, out_ptr),
                    "dsdgpu_decimate_host")

    def submit(self, slot, in_ptr, in_stride, in_first, in_count, fill, p0, nframes, scale, dither_amp, clip,
               dither_first_sample, out_kind, out_ptr):
        self._check(self.lib.dsdgpu_submit(self.h, slot, in_ptr, in_stride, in_first, in_count, fill, p0, nframes,
                                           scale, dither_amp, clip, dither_first_sample, out_kind, out_ptr),
                    "dsdgpu_submit")

    def wait(self, slot):
        self._check(self.lib.dsdgpu_wait(self.h, slot), "dsdgpu_wait")

    def dop_pack_device(self, d_in, in_stride, in_first, in_count, fill, pm0, nframes, reverse_bits, d_out, stream=None):
        self._check(self.lib.dsdgpu_dop_pack_device(self.h, _ptr(d_in), in_stride, in_first, in_count, fill, pm0, nframes,
                                                    int(reverse_bits), _ptr(d_out), stream), "dsdgpu_dop_pack_device")

    def dop_pack_host(self, planar, pm0, nframes, reverse_bits, fill=0x69, in_first=0, in_count=None):
        """planar: uint8 [nch, nbytes] (or the source in the layout set with source_block_bytes)."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        stride = planar.shape[1] if planar.ndim == 2 else 0
        if in_count is None:
            in_count = planar.shape[1]
        out = np.empty((nframes, self.nch), dtype=np.int32)
        self._check(self.lib.dsdgpu_dop_pack_host(self.h, planar.ctypes.data, stride, in_first, in_count, fill, pm0, nframes,
                                                  int(reverse_bits), out.ctypes.data), "dsdgpu_dop_pack_host")
        return out

    def dither_fill_device(self, first_sample, count, d_out, stream=None):
        self._check(self.lib.dsdgpu_dither_fill_device(self.h, first_sample, count, _ptr(d_out), stream),
                    "dsdgpu_dither_fill_device")


def unpack_pcm24(packed):
    """uint8 [..., 3] little-endian two's-complement samples (DSDGPU_OUT_PCM24) -> int32 [...]."""
    p = np.asarray(packed, dtype=np.uint8).astype(np.int32)
    v = p[..., 0] | (p[..., 1] << 8) | (p[..., 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)


class PinnedBuffer:
    """Page-locked host memory from dsdgpu_host_alloc, viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.lib = gpu_lib()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.lib.dsdgpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise DsdGpuError("dsdgpu_host_alloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.dsdgpu_host_free(self.ptr)
            self.ptr = None


class Shard(C.Structure):
    _fields_ = [("job", C.c_int32), ("ch_first", C.c_int32), ("ch_count", C.c_int32), ("reserved", C.c_int32),
                ("frame_first", C.c_int64), ("frame_count", C.c_int64)]


class JobStruct(C.Structure):
    """dsdgpu_job of include/dsdgpu.h."""
    _fields_ = [("in_", _vp), ("in_stride", C.c_size_t), ("in_first", C.c_int64), ("in_count", C.c_int64),
                ("source_block_bytes", C.c_int32), ("fill", C.c_uint8), ("nch", C.c_int32), ("nlut", C.c_int32), ("nstep", C.c_int32),
                ("lut_precision", C.c_int32), ("lut", _vp), ("p0", C.c_int64), ("nframes", C.c_int64), ("scale", C.c_double),
                ("dither_amp", C.c_double), ("clip", C.c_double), ("dither_first_sample", C.c_uint64), ("out_kind", C.c_int32),
                ("out", _vp)]


def run_batch(jobs, devices=(0,), rank=0, world=1, align_frames=1024, lib=None):
    """dsdgpu_run_batch: `jobs` is a list of dicts with the fields of dsdgpu_job (in_/lut/out as addresses or arrays)."""
    lib = lib or gpu_lib()
    arr = (JobStruct * len(jobs))()
    keep = []
    for a, j in zip(arr, jobs):
        for name, _ in JobStruct._fields_:
            v = j[name]
            if name in ("in_", "lut", "out"):
                keep.append(v)
                v = _ptr(v)
            setattr(a, name, v)
    dev = (C.c_int * len(devices))(*devices)
    err = C.create_string_buffer(512)
    rc = lib.dsdgpu_run_batch(arr, len(jobs), dev, len(devices), rank, world, align_frames, err, 512)
    if rc != 0:
        raise DsdGpuError(f"dsdgpu_run_batch failed ({rc}): {err.value.decode()}")


def plan_shards(job_frames, job_nch, job_cost, rank, world, align_frames=1024):
    """Shards of a batch that `rank` of `world` owns (dsdgpu_plan_shards): a list of
    (job, ch_first, ch_count, frame_first, frame_count).  Host arithmetic only, needs no GPU."""
    lib = gpu_lib()
    n = len(job_frames)
    ff = (C.c_int64 * n)(*[int(x) for x in job_frames])
    nc = (C.c_int32 * n)(*[int(x) for x in job_nch])
    jc = (C.c_int32 * n)(*[int(x) for x in job_cost]) if job_cost is not None else None
    cap = sum(int(x) for x in job_nch) + 4
    out = (Shard * cap)()
    got = lib.dsdgpu_plan_shards(ff, nc, jc, n, rank, world, align_frames, out, cap)
    if got < 0 or got > cap:
        raise DsdGpuError(f"dsdgpu_plan_shards failed ({got})")
    return [(o.job, o.ch_first, o.ch_count, o.frame_first, o.frame_count) for o in out[:got]]
