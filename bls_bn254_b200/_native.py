"""Loader / builder of libbls_bn254_b200.so (the C ABI in include/bls_bn254_b200.h).

The library is CUDA-only: there is no CPU implementation behind it.  `lib()` raises if the
shared object is missing, and `Context()` raises if no CUDA device is usable.
"""
import ctypes as C
import os
import pathlib
import shutil
import subprocess

_HERE = pathlib.Path(__file__).resolve().parent
ROOT = _HERE.parent
SO = pathlib.Path(os.environ.get("BN254_SO", str(_HERE / "libbls_bn254_b200.so")))  # BN254_SO: tuning builds
SRC = _HERE / "csrc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC"]

_LIB = None


def sources():
    return sorted(SRC.glob("*.cu")) + sorted(SRC.glob("*.cuh")) + [ROOT / "include" / "bls_bn254_b200.h"]


def needs_build():
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(SO), str(SRC / "capi.cu")]
    subprocess.check_call(cmd)
    return SO


def lib():
    global _LIB
    if _LIB is None:
        if not SO.exists():
            raise RuntimeError(
                f"{SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  bls_bn254_b200 has no CPU fallback.")
        L = C.CDLL(str(SO))
        u8p, u64p, sz, vp = C.c_char_p, C.POINTER(C.c_uint64), C.c_size_t, C.c_void_p
        L.bn254_device_count.restype = C.c_int
        L.bn254_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
        L.bn254_ctx_destroy.argtypes = [vp]
        L.bn254_ctx_destroy.restype = None
        L.bn254_last_error.argtypes = [vp]
        L.bn254_last_error.restype = C.c_char_p
        L.bn254_launch_count.argtypes = [vp]
        L.bn254_launch_count.restype = C.c_uint64
        L.bn254_stream.argtypes = [vp]
        L.bn254_stream.restype = vp
        L.bn254_sync.argtypes = [vp]
        L.bn254_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_long]
        L.bn254_host_register.argtypes = [vp, sz]
        L.bn254_host_unregister.argtypes = [vp]
        L.bn254_comm_unique_id.argtypes = [vp]
        L.bn254_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
        L.bn254_comm_nranks.argtypes = [vp]
        L.bn254_profile_enable.argtypes = [vp, C.c_int]
        L.bn254_profile_drain.argtypes = [vp, vp, sz]
        for name in ("bn254_hash_to_g1", "bn254_encode_to_g1", "bn254_hash_to_g2", "bn254_encode_to_g2"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp, sz, vp, vp]
        for name in ("bn254_pairing_batch", "bn254_miller_loop_batch"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp, vp]
        L.bn254_multi_miller_loop.argtypes = [vp, vp, vp, sz, vp]
        L.bn254_final_exponentiation.argtypes = [vp, vp, sz, vp, vp]
        L.bn254_fp12_product.argtypes = [vp, vp, sz, vp]
        L.bn254_verify_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, sz, vp, sz, C.c_uint32, vp]
        L.bn254_aggregate_verify.argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, sz, C.c_uint32, vp]
        for name in ("bn254_verify_batch_randomized", "bn254_verify_batch_randomized_dev"):
            getattr(L, name).argtypes = [vp, C.c_int, vp, vp, vp, vp, sz, vp, sz, C.c_uint32, vp, C.c_int, sz, vp]
        L.bn254_aggregate_partial.argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, sz, C.c_uint32, vp, vp]
        L.bn254_fast_aggregate_verify_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, sz, vp, sz, C.c_uint32, vp]
        for name in ("bn254_verify_same_key_randomized", "bn254_verify_same_key_randomized_dev"):
            getattr(L, name).argtypes = [vp, C.c_int, vp, vp, sz, vp, vp, vp, vp, sz, C.c_uint32, vp, C.c_int, vp]
        L.bn254_g2_prepare.argtypes = [vp, vp, sz, C.c_uint32, C.POINTER(vp)]
        L.bn254_g2_prepared_free.argtypes = [vp]
        L.bn254_g2_prepared_free.restype = None
        L.bn254_g2_prepared_count.argtypes = [vp]
        L.bn254_g2_prepared_count.restype = sz
        L.bn254_g2_prepared_export.argtypes = [vp, vp, sz, sz, vp, vp]
        for name in ("bn254_verify_batch_prepared", "bn254_verify_batch_prepared_dev"):
            getattr(L, name).argtypes = [vp, vp, vp, vp, vp, vp, sz, vp, sz, C.c_uint32, vp]
        for name in ("bn254_aggregate_verify_dist", "bn254_aggregate_verify_dist_dev"):
            getattr(L, name).argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, sz, C.c_uint32, vp]
        L.bn254_aggregate_finish.argtypes = [vp, vp, vp, sz, vp]
        for name in ("bn254_g1_compress", "bn254_g2_compress", "bn254_g1_is_on_curve", "bn254_g2_is_on_curve", "bn254_g1_is_torsion_free", "bn254_g2_is_torsion_free",
                     "bn254_g1_sum", "bn254_g2_sum"):
            getattr(L, name).argtypes = [vp, vp, sz, vp]
        for name in ("bn254_g1_decompress", "bn254_g2_decompress"):
            getattr(L, name).argtypes = [vp, vp, sz, vp, vp]
        for name in ("bn254_g1_mul", "bn254_g2_mul"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp]
        for name in ("bn254_hash_to_g1_dev", "bn254_hash_to_g2_dev"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp, sz, C.c_int, vp, vp]
        L.bn254_pairing_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp]
        L.bn254_verify_batch_dev.argtypes = [vp, C.c_int, vp, vp, vp, vp, sz, vp, sz, C.c_uint32, vp]
        L.bn254_aggregate_partial_dev.argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, sz, C.c_uint32, vp, vp]
        L.bn254_imad_peak.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.bn254_fp_peak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


EXPORTED_SYMBOLS = [
    "bn254_device_count", "bn254_ctx_create", "bn254_ctx_destroy", "bn254_last_error", "bn254_launch_count", "bn254_stream", "bn254_sync", "bn254_profile_enable", "bn254_profile_drain",
    "bn254_ctx_set_option", "bn254_host_register", "bn254_host_unregister", "bn254_comm_unique_id", "bn254_comm_init_rank", "bn254_comm_nranks",
    "bn254_aggregate_verify_dist", "bn254_aggregate_verify_dist_dev",
    "bn254_fast_aggregate_verify_batch", "bn254_verify_same_key_randomized", "bn254_verify_same_key_randomized_dev",
    "bn254_g2_prepare", "bn254_g2_prepared_free", "bn254_g2_prepared_count", "bn254_g2_prepared_export", "bn254_verify_batch_prepared", "bn254_verify_batch_prepared_dev",
    "bn254_hash_to_g1", "bn254_encode_to_g1", "bn254_hash_to_g2", "bn254_encode_to_g2",
    "bn254_pairing_batch", "bn254_miller_loop_batch", "bn254_multi_miller_loop", "bn254_final_exponentiation", "bn254_fp12_product",
    "bn254_verify_batch", "bn254_verify_batch_randomized", "bn254_aggregate_verify", "bn254_aggregate_partial", "bn254_aggregate_finish",
    "bn254_g1_compress", "bn254_g1_decompress", "bn254_g2_compress", "bn254_g2_decompress", "bn254_g1_is_on_curve", "bn254_g2_is_on_curve",
    "bn254_g1_is_torsion_free", "bn254_g2_is_torsion_free",
    "bn254_g1_mul", "bn254_g2_mul", "bn254_g1_sum", "bn254_g2_sum",
    "bn254_hash_to_g1_dev", "bn254_hash_to_g2_dev", "bn254_pairing_batch_dev", "bn254_verify_batch_dev", "bn254_verify_batch_randomized_dev", "bn254_aggregate_partial_dev",
    "bn254_imad_peak", "bn254_fp_peak",
]
