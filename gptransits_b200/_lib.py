"""ctypes binding of the C ABI in include/gptransits_b200.h.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded the
import of the product path fails loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPT_LIB_PATH") or os.path.join(_HERE, "libgptransits_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("engine.cu", "kernels.cuh", "device_math.cuh")]
HEADER = os.path.join(os.path.dirname(_HERE), "include", "gptransits_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DTR_MINBLOCKS=3",
              "--shared", "-Xcompiler", "-fPIC"]


class GptError(RuntimeError):
    pass


class Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class ModelDesc(C.Structure):
    _fields_ = [("ncomp", C.c_int32), ("comp_type", C.POINTER(C.c_int32)), ("has_transit", C.c_int32),
                ("tr_fixed", C.c_double * 9), ("tr_free", C.c_uint8 * 9), ("supersample", C.c_int32),
                ("exp_time_days", C.c_double), ("ellip_mode", C.c_int32), ("inc_mode", C.c_int32), ("ndim", C.c_int32),
                ("priors", C.POINTER(Prior))]


def source_hash():
    """Hash of everything the library is built from; embedded in the binary (gpt_version)."""
    import hashlib
    h = hashlib.sha256()
    for path in SOURCES + [HEADER]:
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _built_hash(path):
    """The source hash embedded in a built library, read from the file (no dlopen)."""
    try:
        with open(path, "rb") as f:
            data = f.read()
        i = data.find(b"(sm_100a, src=")
        if i < 0:
            return None
        return data[i + 14:i + 30].decode()
    except Exception:
        return None


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree with nvcc for sm_100a."""
    want = source_hash()
    if not force and os.path.exists(LIB_PATH) and _built_hash(LIB_PATH) == want:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f'-DGPT_SRC_HASH="{want}"', "-o", LIB_PATH, SOURCES[0]]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/gptransits_b200.h
SIGNATURES = {
    "gpt_engine_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gpt_engine_destroy": (None, [_vp]),
    "gpt_last_error": (C.c_char_p, [_vp]),
    "gpt_version": (C.c_char_p, []),
    "gpt_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "gpt_get_stat": (C.c_int, [_vp, C.c_char_p, _dp]),
    "gpt_model_set": (C.c_int, [_vp, C.POINTER(ModelDesc)]),
    "gpt_star_upload": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int64]),
    "gpt_stars_upload": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int64]),
    "gpt_star_clear": (C.c_int, [_vp]),
    "gpt_log_prob": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp, _dp]),
    "gpt_log_prob_multi": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp]),
    "gpt_log_prob_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp]),
    "gpt_transit_flux": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp]),
    "gpt_gp_predict": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int64, _dp, _dp]),
    "gpt_sample": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_uint64, C.c_uint64,
                             C.c_double, _dp, _dp, _i64p]),
    "gpt_sampler_begin": (C.c_int, [_vp, C.c_int, C.c_int, _i32p, _dp, C.c_int, C.c_uint64, C.c_uint64,
                                    C.c_double]),
    "gpt_sampler_run": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "gpt_sampler_read": (C.c_int, [_vp, C.c_int64, C.c_int64, _dp, _dp]),
    "gpt_sampler_state": (C.c_int, [_vp, _dp, _dp, _i64p]),
    "gpt_sampler_rhat": (C.c_int, [_vp, C.c_int64, C.c_int64, _dp]),
    "gpt_sampler_geweke": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int, _i64p, C.c_double, C.c_double, _dp]),
    "gpt_sampler_end": (C.c_int, [_vp]),
    "gpt_sampler_reserve": (C.c_int, [_vp, C.c_int64]),
    "gpt_sampler_lq_device": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "gpt_sampler_half_eval": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "gpt_sampler_half_accept": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gpt_comm_id": (C.c_int, [_vp]),
    "gpt_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "gpt_comm_destroy": (C.c_int, [_vp]),
    "gpt_sampler_run_sharded": (C.c_int, [_vp, C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "gpt_measure_fp64_peak": (C.c_int, [_vp, _dp]),
    "gpt_sampler_run_timed": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "gpt_log_prob_device_timed": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                            C.POINTER(C.c_float)]),
    "gpt_device_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "gpt_device_free": (C.c_int, [_vp, _vp]),
    "gpt_device_copy": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int]),
}


def lib():
    """Load libgptransits_b200.so (built by __graft_entry__.build()); raise if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GptError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.gpt_version.restype = C.c_char_p
        built = _lib.gpt_version().decode()
        if "GPT_LIB_PATH" not in os.environ and f"src={source_hash()}" not in built:
            _lib = None
            raise GptError(f"{LIB_PATH} is stale ({built}; sources now hash to {source_hash()}): rebuild with "
                           "`python -c 'import __graft_entry__ as g; g.build()'`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib
