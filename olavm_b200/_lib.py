"""ctypes loader of libola_gpu.so (the C ABI declared in include/ola_gpu.h).

The product path has NO fallback: if the shared library is missing, or no sm_100 device is visible,
`load()` / `Context()` raise.  Nothing here imports the oracle.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libola_gpu.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ola_gpu.h")

OLA_OK = 0
ERRORS = {
    -1: "OLA_ERR_NO_DEVICE",
    -2: "OLA_ERR_CUDA",
    -3: "OLA_ERR_INVALID_ARG",
    -4: "OLA_ERR_OOM",
    -5: "OLA_ERR_QUOTIENT_DEGREE",
    -6: "OLA_ERR_ZETA_IN_SUBGROUP",
    -7: "OLA_ERR_INTERNAL",
}


class OlaError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")


u64p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/ola_gpu.h one to one (tests check the two agree)
SIGNATURES = {
    "ola_gpu_init": (_int, [_int, ctypes.POINTER(_vp)]),
    "ola_gpu_destroy": (None, [_vp]),
    "ola_gpu_last_error": (ctypes.c_char_p, [_vp]),
    "ola_gpu_sync": (_int, [_vp]),
    "ola_gpu_kernel_launches": (_u64, [_vp]),
    "ola_gpu_stream": (_vp, [_vp]),
    "ola_profile_begin": (_int, [_vp]),
    "ola_profile_end": (_int, [_vp, ctypes.c_char_p, _sz]),
    "ola_dev_alloc": (_int, [_vp, _sz, ctypes.POINTER(_vp)]),
    "ola_dev_free": (_int, [_vp, _vp]),
    "ola_dev_upload": (_int, [_vp, _vp, _vp, _sz]),
    "ola_dev_download": (_int, [_vp, _vp, _vp, _sz]),
    "ola_dev_copy": (_int, [_vp, _vp, _vp, _sz]),
    "ola_dev_gather_rows": (_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "ola_ntt_forward": (_int, [_vp, _vp, _int, _sz, _u32]),
    "ola_ntt_inverse": (_int, [_vp, _vp, _int, _sz, _u32]),
    "ola_coset_lde": (_int, [_vp, _vp, _vp, _int, _sz, _u32, _u32, _u64, _int]),
    "ola_coset_intt": (_int, [_vp, _vp, _int, _sz, _u32, _u64]),
    "ola_poseidon_permute": (_int, [_vp, _vp, _int, _sz]),
    "ola_hash_rows": (_int, [_vp, _vp, _vp, _int, _sz, _sz]),
    "ola_merkle_rows": (_int, [_vp, _vp, _int, _sz, _sz, _u32, _vp, _vp]),
    "ola_lde_batch": (_int, [_vp, _vp, _int, _sz, _u32, _int, _u32, _vp, _vp]),
    "ola_commit": (_int, [_vp, _vp, _int, _sz, _u32, _int, _u32, _u32, ctypes.POINTER(_vp), _vp]),
    "ola_commit_shard": (_int, [_vp, _vp, _int, _sz, _u32, _int, _u32, _u32, _u32, _u32, ctypes.POINTER(_vp), _vp]),
    "ola_prove": (_int, [_vp, ctypes.POINTER(_int), _u32, ctypes.POINTER(_vp), _int, ctypes.POINTER(_u32), _vp, _int, _vp, _sz,
                         ctypes.POINTER(_sz)]),
    "ola_table_columns": (_int, [_int]),
    "ola_air_constraints": (_int, [_int, _vp, _vp, _u64, _vp, _vp, _int]),
    "ola_generate_poseidon_trace": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _int]),
    "ola_permuted_cols": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _int]),
    "ola_generate_rangecheck_trace": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_bitwise_trace": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _u32, _vp, ctypes.POINTER(_u64), _int]),
    "ola_generate_cmp_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_cpu_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_memory_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_program_trace": (_int, [_vp, _vp, _sz, _vp, _sz, _vp, _u32, _vp, ctypes.POINTER(_u64), _int]),
    "ola_generate_poseidon_chunk_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_storage_access_trace": (_int, [_vp, _vp, _sz, _sz, _u32, _vp, _int]),
    "ola_generate_tape_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_sccall_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_generate_prog_chunk_trace": (_int, [_vp, _vp, _sz, _u32, _vp, _int]),
    "ola_trace_from_json": (_int, [ctypes.c_char_p, _sz, ctypes.POINTER(_vp), ctypes.c_char_p, _sz]),
    "ola_trace_free": (None, [_vp]),
    "ola_trace_new": (_int, [ctypes.POINTER(_vp)]),
    "ola_trace_set_records": (_int, [_vp, _int, _vp, _sz]),
    "ola_trace_records": (_int, [_vp, _int, ctypes.POINTER(_vp), ctypes.POINTER(_sz), ctypes.POINTER(_u32)]),
    "ola_trace_table_log_rows": (_int, [_vp, _int]),
    "ola_generate_traces": (_int, [_vp, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_u32), ctypes.POINTER(_u64)]),
    "ola_prove_trace": (_int, [_vp, _vp, _vp, _sz, ctypes.POINTER(_sz)]),
    "ola_compress_challenge": (_int, [ctypes.POINTER(_vp), _u32, _sz, ctypes.POINTER(_u64)]),
    "ola_verify": (_int, [ctypes.POINTER(_int), _u32, _vp, _sz, ctypes.c_char_p, _sz]),
    "ola_verify_cfg": (_int, [_int, ctypes.POINTER(_int), _u32, _vp, _sz, ctypes.c_char_p, _sz]),
    "ola_verify_subsystem_cfg": (_int, [_int, ctypes.POINTER(_int), _u32, _vp, _sz, ctypes.c_char_p, _sz]),
    "ola_set_hasher": (_int, [_vp, _int]),
    "ola_get_hasher": (_int, [_vp]),
    "ola_set_comm": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
    "ola_prove_session_begin": (_int, [_vp, ctypes.POINTER(_int), _u32, ctypes.POINTER(_vp), _int, ctypes.POINTER(_u32), _vp, _int, ctypes.POINTER(_vp)]),
    "ola_prove_session_next": (_int, [_vp, _vp]),
    "ola_prove_session_supply": (_int, [_vp, _vp, _sz]),
    "ola_prove_session_finish": (_int, [_vp, _vp, _sz, ctypes.POINTER(_sz)]),
    "ola_nccl_unique_id": (_int, [ctypes.c_char_p, _vp]),
    "ola_set_comm_nccl": (_int, [_vp, ctypes.c_char_p, _int, _int, _vp]),
    "ola_comm_bytes": (_u64, [_vp]),
    "ola_batch_free": (_int, [_vp, _vp]),
    "ola_batch_ncols": (_sz, [_vp]),
    "ola_batch_degree_log": (_u32, [_vp]),
    "ola_batch_rate_bits": (_u32, [_vp]),
    "ola_batch_coeffs_dev": (_vp, [_vp]),
    "ola_batch_lde_dev": (_vp, [_vp]),
    "ola_batch_nodes_dev": (_vp, [_vp]),
    "ola_batch_get_coeffs": (_int, [_vp, _vp, _vp]),
    "ola_batch_get_cap": (_int, [_vp, _vp, _vp]),
    "ola_batch_get_leaves": (_int, [_vp, _vp, _sz, _sz, _vp]),
    "ola_batch_prove_leaf": (_int, [_vp, _vp, _sz, _vp]),
}


def header_symbols():
    """Function names declared in include/ola_gpu.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ola_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load():
    """dlopen libola_gpu.so and bind every symbol of the header.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OlaError(-7, f"{SO_PATH} not built: run `python -m olavm_b200.build` (there is no CPU fallback)")
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def hptr(a):
    """Host numpy uint64 array -> void* (must be C-contiguous)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need contiguous uint64"
    return a.ctypes.data_as(_vp)
