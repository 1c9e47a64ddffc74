"""ctypes loader for liblva_b200.so (include/lva_b200.h).  No fallback: if the CUDA library has
not been built, importing a decoder fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# LVA_B200_LIB: developer override to A/B two builds of the same CUDA library (never a CPU path)
LIB_PATH = os.environ.get("LVA_B200_LIB") or os.path.join(HERE, "_build", "liblva_b200.so")
CLI_PATH = os.path.join(HERE, "_build", "viterbi_nanopore.out")

LVA_DEFAULT_DEVIATION = 0xFFFFFFFF
LVA_FLAG_EXACT_ONLY = 1
LVA_FLAG_UNFUSED_BAND = 2

EXPORTS = ["lva_create", "lva_destroy", "lva_last_error", "lva_decode_batch", "lva_decode_batch_device", "lva_decode_batch_crc", "lva_decode_ranges",
           "lva_trim_batch", "lva_posteriors_batch", "lva_posteriors_batch_device", "lva_encode", "lva_check_params", "lva_get_code_tables", "lva_band", "lva_band_ex", "lva_nstate_pos", "lva_slots", "lva_count_work", "lva_get_timing", "lva_device_count", "lva_version"]


class LvaParams(C.Structure):
    _fields_ = [("mem_conv", C.c_uint32), ("rate", C.c_uint32), ("msg_len", C.c_uint32),
                ("list_size", C.c_uint32), ("max_deviation", C.c_uint32), ("rc", C.c_int32),
                ("sync_marker", C.c_char_p), ("sync_period", C.c_uint32), ("device", C.c_int32),
                ("max_slots", C.c_uint32), ("flags", C.c_uint32)]


class LvaCodeTables(C.Structure):
    _fields_ = [("G", C.c_uint32 * 2), ("initial_state", C.c_uint32), ("final_state", C.c_uint32),
                ("punc_len", C.c_uint32), ("nstate_pos", C.c_uint32), ("punc", C.c_uint8 * 16),
                ("pos2msg", C.c_uint32 * 260), ("valid_mask", C.c_uint32 * 260),
                ("valid_value", C.c_uint32 * 260), ("history_words", C.c_uint8 * 260),
                ("first_run_row", C.c_uint8 * 260)]


class LvaWork(C.Structure):
    _fields_ = [("state_updates", C.c_double), ("edge_acs", C.c_double), ("state_bytes", C.c_double)]


class LvaTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("kernel_ms", C.c_float), ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float), ("launches", C.c_uint64), ("step_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("exact_fallbacks", C.c_uint64),
                ("chain_inserts", C.c_uint64), ("chain_keys_max", C.c_uint64),
                ("chain_table_slots", C.c_uint64), ("table_regrows", C.c_uint64),
                ("structure_check_failures", C.c_uint64)]


def build(verbose: bool = False) -> None:
    """Compile the CUDA library and CLI in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the decode path)")
    L = C.CDLL(LIB_PATH)
    L.lva_create.restype = C.c_void_p
    L.lva_create.argtypes = [C.POINTER(LvaParams), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    L.lva_destroy.argtypes = [C.c_void_p]
    L.lva_last_error.restype = C.c_char_p
    L.lva_last_error.argtypes = [C.c_void_p]
    for fn in (L.lva_decode_batch, L.lva_decode_batch_device):
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p]
    L.lva_decode_ranges.restype = C.c_int
    L.lva_decode_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    L.lva_decode_batch_crc.restype = C.c_int
    L.lva_decode_batch_crc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    L.lva_trim_batch.restype = C.c_int
    L.lva_trim_batch.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p,
                                 C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for fn in (L.lva_posteriors_batch, L.lva_posteriors_batch_device):
        fn.restype = C.c_int
        fn.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.lva_encode.restype = C.c_int
    L.lva_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                             C.POINTER(C.c_uint32)]
    L.lva_nstate_pos.restype = C.c_uint32
    L.lva_nstate_pos.argtypes = [C.c_void_p]
    L.lva_slots.restype = C.c_uint32
    L.lva_slots.argtypes = [C.c_void_p]
    L.lva_count_work.restype = C.c_int
    L.lva_count_work.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(LvaWork)]
    L.lva_get_timing.restype = C.c_int
    L.lva_get_timing.argtypes = [C.c_void_p, C.POINTER(LvaTiming)]
    L.lva_get_code_tables.restype = C.c_int
    L.lva_get_code_tables.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_char_p, C.c_uint32,
                                      C.POINTER(LvaCodeTables)]
    L.lva_band.restype = None
    L.lva_band.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)] * 2
    L.lva_band_ex.restype = None
    L.lva_band_ex.argtypes = [C.c_uint32] * 5 + [C.POINTER(C.c_uint32)] * 2
    L.lva_check_params.restype = C.c_int
    L.lva_check_params.argtypes = [C.c_uint32] * 3
    L.lva_version.restype = C.c_char_p
    _lib = L
    return L
