"""ctypes binding of libsylber_b200.so (the C ABI in include/sylber_b200.h) and its nvcc build recipe."""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_SO = os.path.join(_PKG, "libsylber_b200.so")

SYL_SPLIT_CONV, SYL_SPLIT_PROJ, SYL_SPLIT_ENC, SYL_SPLIT_CONV1 = 1, 2, 4, 8
SYL_SPLIT_CONV2, SYL_SPLIT_CONV3, SYL_SPLIT_CONV4, SYL_SPLIT_CONV5, SYL_SPLIT_CONV6 = 16, 32, 64, 128, 256
SYL_SPLIT_FPROJ, SYL_SPLIT_POS = 512, 1024
SYL_TRIM_PADDING = 4096
MODES = {
    "parity": SYL_SPLIT_CONV4 | SYL_SPLIT_CONV5 | SYL_SPLIT_CONV6 | SYL_SPLIT_FPROJ,
    "strict": SYL_SPLIT_CONV | SYL_SPLIT_CONV1 | SYL_SPLIT_PROJ,
    "fast": 0,
    "exact": SYL_SPLIT_CONV | SYL_SPLIT_CONV1 | SYL_SPLIT_PROJ | SYL_SPLIT_ENC,
}

_c_void_p, _c_int, _c_size_t, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "syl_create": (_c_int, [ctypes.POINTER(_c_void_p), _c_int, _c_int, _c_int]),
    "syl_load_weight": (_c_int, [_c_void_p, ctypes.c_char_p, _c_void_p, ctypes.POINTER(ctypes.c_int64), _c_int, _c_int]),
    "syl_finalize": (_c_int, [_c_void_p]),
    "syl_destroy": (None, [_c_void_p]),
    "syl_last_error": (ctypes.c_char_p, [_c_void_p]),
    "syl_num_frames": (_c_int, [_c_int]),
    "syl_workspace_bytes": (_c_size_t, [_c_void_p, _c_int, _c_int]),
    "syl_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                             _c_void_p, _c_int, _c_float, _c_float, _c_void_p, _c_size_t, _c_void_p]),
    "syl_conv_frontend": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "syl_encoder_layer": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p,
                                   _c_size_t, _c_void_p]),
    "syl_attention": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "syl_pcm16_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "syl_prepare_pcm16": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "syl_prepare_f32": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "syl_resample": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_int,
                              _c_void_p]),
    "syl_kmeans_assign": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "syl_segment_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "syl_segment": (_c_int, [_c_void_p, _c_int, _c_int, _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_int,
                             _c_void_p, _c_size_t, _c_void_p]),
    "syl_gemm_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "syl_gemm_f32": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                              _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "syl_powf_half": (_c_int, [_c_void_p, _c_void_p, ctypes.c_int64, _c_void_p]),
    "syl_read_stage": (_c_int, [_c_void_p, ctypes.c_char_p, _c_void_p, _c_size_t, _c_void_p]),
    "syl_saturation_scan": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "syl_set_active_layers": (_c_int, [_c_void_p, _c_int]),
    "syl_forward_launch_count": (_c_int, [_c_void_p, _c_int]),
    "syl_set_graph_mode": (_c_int, [_c_void_p, _c_int]),
    "syl_num_stages": (_c_int, []),
    "syl_stage_name": (ctypes.c_char_p, [_c_int]),
    "syl_profile_enable": (_c_int, [_c_void_p, _c_int]),
    "syl_profile_read": (_c_int, [_c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
}

# entry points that exist only in the diagnostic build (-DSYL_DIAG): tools/attn_trace.py, tools/mma_probe.py
DIAG_SIGNATURES = {
    "syl_attention_trace": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "syl_mma_probe": (_c_int, [_c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "syl_gemm_set_trace": (_c_int, [_c_void_p]),
}
_SO_DIAG = os.path.join(_PKG, "libsylber_b200_diag.so")

_LIB = None
_LIB_DIAG = None


def library_path() -> str:
    return _SO


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(os.path.dirname(_PKG), "include", "sylber_b200.h")]


def build_library(force: bool = False, verbose: bool = False, diag: bool = False, defines=(), out: str = None) -> str:
    """Compile csrc/api.cu for sm_100a into sylber_b200/libsylber_b200.so (in-tree so it travels to the GPU box).
    `diag=True` builds libsylber_b200_diag.so with -DSYL_DIAG instead: the SYL_* environment switches and the
    diagnostic entry points (DIAG_SIGNATURES), which the product library does not contain.  `defines` / `out` build a
    variant for an A/B measurement into another file (tools/ab_lib.sh swaps it in on the GPU box)."""
    _SO = out or (_SO_DIAG if diag else globals()["_SO"])
    if not force and os.path.exists(_SO):
        newest = max(os.path.getmtime(s) for s in _sources())
        if os.path.getmtime(_SO) >= newest:
            return _SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsylber_b200.so")
    tmp = f"{_SO}.{os.getpid()}.tmp"       # never leave a half-written library where another process may load it
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
           "-o", tmp, os.path.join(_CSRC, "api.cu")]
    if diag:
        cmd.insert(1, "-DSYL_DIAG")
    for d in defines:
        cmd.insert(1, "-D" + d)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    try:
        subprocess.check_call(cmd, cwd=_CSRC)
        os.replace(tmp, _SO)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return _SO


def load_library():
    """Load the C-ABI library.  Raises if it has not been built - there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise RuntimeError(
            f"{_SO} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); sylber_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(_SO)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def load_diag_library():
    """The -DSYL_DIAG build (experiment switches from the environment + diagnostic entry points); tools only."""
    global _LIB_DIAG
    if _LIB_DIAG is None:
        lib = ctypes.CDLL(build_library(diag=True))
        for name, (res, args) in {**SIGNATURES, **DIAG_SIGNATURES}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB_DIAG = lib
    return _LIB_DIAG


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.syl_last_error(handle)
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")
