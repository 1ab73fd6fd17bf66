"""ctypes binding of csrc/libnsagp.so (the C ABI declared in include/nsagp.h).

There is no CPU implementation behind these calls: if the shared library is
missing the import fails loudly, and on a machine without a CUDA device every
compute call raises ``NsagpError`` (status NSAGP_ERR_CUDA).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

# Kernels must be loaded when the library is: a rank that spins in comm_exchange_kernel for a peer's record must never
# keep that peer (another thread of the same process in the emulated-rank tests, csrc/comm.cuh) from loading a kernel
# lazily -- a lazy load synchronises with the running spin kernel and the two wait for each other.  Read by the driver
# at CUDA initialisation, so it only helps if set before that; bench.py and tests/conftest.py set it first thing too.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# NSAGP_LIB: load another build of the same sources (e.g. an instrumented debug build)
LIB_PATH = os.environ.get("NSAGP_LIB") or os.path.join(CSRC, "libnsagp.so")
_SOURCES = ["api.cu", "api_full.inc", "common.cuh", "mom.cuh", "momcta.cuh", "mombatch.cuh", "lookup.cuh", "ihgp.cuh",
            "gfep.cuh", "adfcta.cuh", "fastmath.cuh", "scan.cuh", "ekf.cuh", "ekfscan.cuh", "mcrec.cuh", "api_mc.inc", "api_ekf.inc", "api_chunk.inc", "api_tables.inc", "comm.cuh",
            "api_comm.inc", "siteupd.cuh", "ekfbig.cuh", "fastfb.cuh", "api_fb.inc", "ekfgrad.cuh", "logtab.inc"]

c_double_p = C.POINTER(C.c_double)


class NsagpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("nsagp status %d: %s" % (status, message))
        self.status = status


class Model(C.Structure):
    _fields_ = [("D", C.c_int32), ("N", C.c_int32), ("bz", C.c_int32), ("bg", C.c_int32),
                ("A", c_double_p), ("Q", c_double_p), ("Pinf", c_double_p), ("h", c_double_p)]


class Lik(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sn2", C.c_double), ("link_shift", C.c_double), ("W", c_double_p),
                ("S", C.c_int32), ("wn", c_double_p), ("xn", c_double_p)]


class Ep(C.Structure):
    _fields_ = [("ep_fraction", C.c_double), ("ep_damping", c_double_p), ("ep_itts", C.c_int32)]


class Tables(C.Structure):
    _fields_ = [("nr", C.c_int32), ("r", c_double_p), ("PP", c_double_p), ("PG", c_double_p)]


_OUT_FIELDS = ["Eft", "Varft", "lb", "ub", "ttau", "tnu", "R", "lZ", "MF", "MS", "PF", "PS",
               "nlZ", "maxDiffM", "maxDiffP", "edata"]


class Outputs(C.Structure):
    _fields_ = [(f, c_double_p) for f in _OUT_FIELDS] + [("n_negcav", C.POINTER(C.c_int64))]


MODE_PREDICT, MODE_NLZ, MODE_NLZ_RUNNING = 0, 1, 2
KIND_IHGP, KIND_FULL = 0, 1


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a into csrc/libnsagp.so (in-tree, so it travels to
    the GPU box).  nvcc cross-compiles without a GPU."""
    srcs = [os.path.join(CSRC, s) for s in _SOURCES] + [os.path.join(_HERE, "..", "include", "nsagp.h")]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs if os.path.exists(s))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # NSAGP_FAST_BUILD=1 adds --split-compile=0 (back end in parallel on all host cores: 4.5 min -> 1.5 min) for
    # development only: the split build measured 7-17 % slower kernels on the B200 (bench.py phases), so the shipped
    # library is always the serial compile.
    fast = ["--split-compile=0"] if os.environ.get("NSAGP_FAST_BUILD") == "1" else []
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"] + fast + [
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", LIB_PATH,
           os.path.join(CSRC, "api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.nsagp_version.restype = C.c_char_p
    L.nsagp_last_error.restype = C.c_char_p
    L.nsagp_launch_count.restype = C.c_int64
    L.nsagp_launch_count.argtypes = [C.c_int]
    L.nsagp_set_device.argtypes = [C.c_int]
    L.nsagp_set_stream.argtypes = [C.c_void_p]
    mom_args = [C.POINTER(Lik), C.c_int32, C.c_int32, C.c_double, C.c_int64, c_double_p, c_double_p, c_double_p,
                c_double_p, c_double_p, c_double_p]
    L.nsagp_mom_batch.argtypes = mom_args
    L.nsagp_mom_batch_warp.argtypes = mom_args
    L.nsagp_ep_ihgp.argtypes = [C.POINTER(Model), C.POINTER(Lik), C.POINTER(Ep), C.POINTER(Tables), c_double_p,
                                C.c_int64, C.c_int32, C.POINTER(Outputs)]
    L.nsagp_ep_full.argtypes = [C.POINTER(Model), C.POINTER(Lik), C.POINTER(Ep), c_double_p, C.c_int64, C.c_int32,
                                C.POINTER(Outputs)]
    L.nsagp_ep_ihgp_batch.argtypes = [C.c_int32] + L.nsagp_ep_ihgp.argtypes
    L.nsagp_ep_full_batch.argtypes = [C.c_int32] + L.nsagp_ep_full.argtypes
    L.nsagp_plan_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.POINTER(Model), C.POINTER(Lik),
                                    C.POINTER(Ep), C.POINTER(Tables), c_double_p, C.c_int64, C.c_int32]
    L.nsagp_plan_run.argtypes = [C.c_void_p]
    L.nsagp_plan_fetch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Outputs)]
    L.nsagp_plan_destroy.argtypes = [C.c_void_p]
    L.nsagp_plan_timings.argtypes = [C.c_void_p, c_double_p, C.c_int32]
    L.nsagp_plan_keep_pf.argtypes = [C.c_void_p, C.c_int]
    L.nsagp_plan_set_adf_form.argtypes = [C.c_void_p, C.c_int]
    L.nsagp_giekf.argtypes = [C.POINTER(Model), c_double_p, C.c_double, C.c_int32, C.c_int32, c_double_p, C.c_int64,
                              C.c_int32, C.POINTER(Outputs)]
    L.nsagp_mc_reconstruct.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p, C.c_double,
                                       C.c_int32, c_double_p, C.c_uint64, c_double_p, c_double_p, c_double_p, c_double_p]
    L.nsagp_giekf_carry.argtypes = L.nsagp_giekf.argtypes
    L.nsagp_giekf_grad.argtypes = [C.POINTER(Model), c_double_p, C.c_double, C.c_int32, C.POINTER(C.c_int32), c_double_p,
                                   c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]
    L.nsagp_giekf_config.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.nsagp_giekf_timings.argtypes = [c_double_p, C.c_int32]
    L.nsagp_plan_set_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.nsagp_plan_stage.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_int64]
    L.nsagp_ihgp_tables.argtypes = [C.POINTER(Model), C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, c_double_p,
                                    c_double_p, c_double_p]
    L.nsagp_fastmath_eval.argtypes = [C.c_int32, C.c_int64, c_double_p, c_double_p]
    L.nsagp_fastfb.argtypes = [C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, C.c_int64,
                               c_double_p, c_double_p]
    L.nsagp_scan_config.argtypes = [C.c_int64]
    L.nsagp_scan_merge.argtypes = [C.c_int32]
    L.nsagp_scan_prefetch.argtypes = [C.c_int32]
    L.nsagp_scan_tile.argtypes = [C.c_int32, C.c_int32]
    L.nsagp_site_config.argtypes = [C.c_int32]
    L.nsagp_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64]
    L.nsagp_comm_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.nsagp_comm_connect.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.nsagp_comm_destroy.argtypes = [C.c_void_p]
    L.nsagp_plan_comm_slot_doubles.argtypes = [C.c_void_p]
    L.nsagp_plan_comm_slot_doubles.restype = C.c_int64
    L.nsagp_plan_run_chunked.argtypes = [C.c_void_p, C.c_void_p]
    L.nsagp_plan_set_adf_parallel.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
    L.nsagp_plan_adf_mismatch.argtypes = [C.c_void_p, c_double_p]
    _lib = L
    return L


EXPORTS = ["nsagp_version", "nsagp_last_error", "nsagp_set_device", "nsagp_set_stream", "nsagp_device_count",
           "nsagp_launch_count", "nsagp_mom_batch", "nsagp_mom_batch_warp", "nsagp_ep_ihgp", "nsagp_ep_full",
           "nsagp_ep_ihgp_batch", "nsagp_ep_full_batch", "nsagp_plan_create", "nsagp_plan_run",
           "nsagp_plan_fetch", "nsagp_plan_destroy", "nsagp_plan_timings", "nsagp_plan_keep_pf",
           "nsagp_plan_set_adf_form", "nsagp_fastmath_eval", "nsagp_release_cache", "nsagp_giekf", "nsagp_plan_set_range",
           "nsagp_plan_stage", "nsagp_ihgp_tables", "nsagp_giekf_config", "nsagp_giekf_timings", "nsagp_giekf_carry", "nsagp_mc_reconstruct",
           "nsagp_comm_create", "nsagp_comm_export", "nsagp_comm_connect", "nsagp_comm_destroy", "nsagp_plan_comm_slot_doubles",
           "nsagp_plan_run_chunked", "nsagp_plan_set_adf_parallel", "nsagp_plan_adf_mismatch", "nsagp_fastfb", "nsagp_scan_config", "nsagp_scan_merge", "nsagp_scan_prefetch", "nsagp_scan_tile", "nsagp_site_config", "nsagp_giekf_grad"]


def check(status):
    if status != 0:
        raise NsagpError(status, lib().nsagp_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))
