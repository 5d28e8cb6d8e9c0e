"""ctypes binding of libnucleo_b200.so (include/nucleo_b200.h).  No torch, no CPU fallback.

The library is built in-tree by ``python -m nucleoatac_b200.build`` (nvcc, sm_100a).  Loading
fails loudly when the shared object is missing, and ``nb200_ctx_create`` fails loudly when
there is no B200 -- nothing in this package computes on the CPU instead.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NB200_LIB") or os.path.join(HERE, "libnucleo_b200.so")  # NB200_LIB: developer override (kernel variants)

c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_double_p = C.POINTER(C.c_double)
c_uint8_p = C.POINTER(C.c_uint8)


class NB200Error(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, msg)
        self.code = code


OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_CAPACITY, ERR_FLANK = range(6)


class OccParams(C.Structure):
    _fields_ = [("upper", C.c_int32), ("flank", C.c_int32), ("step", C.c_int32), ("sep", C.c_int32),
                ("min_occ", C.c_double), ("atac", C.c_int32), ("use_bias", C.c_int32),
                ("smooth_win", c_double_p), ("smooth_len", C.c_int32)]


class NucParams(C.Structure):
    _fields_ = [("atac", C.c_int32), ("use_bias", C.c_int32), ("smooth_sd", C.c_int32),
                ("nonredundant_sep", C.c_int32), ("redundant_sep", C.c_int32),
                ("min_z", C.c_double), ("min_lr", C.c_double), ("min_reads", C.c_double),
                ("smooth_win", c_double_p), ("smooth_len", C.c_int32), ("xcor_mode", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("n_chunks", C.c_int32), ("chunk_start", c_int32_p), ("chunk_end", c_int32_p),
                ("frag_off", c_int64_p), ("frag_pos", c_int32_p), ("frag_tlen", c_int32_p),
                ("seq_off", c_int64_p), ("seq_start", c_int32_p), ("seq", c_uint8_p)]


class OccOut(C.Structure):
    _fields_ = [("smoothed_vals", c_double_p), ("smoothed_lower", c_double_p), ("smoothed_upper", c_double_p),
                ("vals", c_double_p), ("lower_bound", c_double_p), ("upper_bound", c_double_p),
                ("cov", c_double_p), ("nuc_dist", c_double_p), ("peak_count", c_int32_p),
                ("peak_off", c_int64_p), ("peak_pos", c_int32_p), ("peak_occ", c_double_p),
                ("peak_lower", c_double_p), ("peak_upper", c_double_p), ("peak_reads", c_double_p)]


class NucOut(C.Structure):
    _fields_ = [("nuc_signal", c_double_p), ("background", c_double_p), ("norm_signal", c_double_p),
                ("smoothed", c_double_p), ("nuc_cov", c_double_p), ("nfr_cov", c_double_p),
                ("cand_count", c_int32_p), ("cand_off", c_int64_p), ("cand_pos", c_int32_p),
                ("cand_flag", c_int32_p), ("cand_z", c_double_p), ("cand_lr", c_double_p),
                ("cand_norm_signal", c_double_p), ("cand_nuc_signal", c_double_p), ("cand_nuc_cov", c_double_p),
                ("cand_nfr_cov", c_double_p), ("cand_smoothed", c_double_p)]


c_float_p = C.POINTER(C.c_float)


class OccOut32(C.Structure):   # nb200_occ_out32: the per-position tracks as float32
    _fields_ = [(n, c_float_p if i < 7 else t) for i, (n, t) in enumerate(OccOut._fields_)]


class NucOut32(C.Structure):   # nb200_nuc_out32
    _fields_ = [(n, c_float_p if i < 6 else t) for i, (n, t) in enumerate(NucOut._fields_)]


_lib = None


def _sig(lib, name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


def load():
    """dlopen libnucleo_b200.so and declare every entry point of include/nucleo_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libnucleo_b200.so is not built (run `python -m nucleoatac_b200.build`); "
                          "nucleoatac_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    I = C.c_int
    _sig(lib, "nb200_ctx_create", I, I, C.POINTER(vp))
    _sig(lib, "nb200_ctx_destroy", I, vp)
    _sig(lib, "nb200_last_error", C.c_char_p, vp)
    _sig(lib, "nb200_device_info", I, vp, C.POINTER(I), C.POINTER(I), C.POINTER(I), c_int64_p)
    _sig(lib, "nb200_host_alloc", I, vp, i64, C.POINTER(vp))
    _sig(lib, "nb200_host_free", I, vp, vp)
    _sig(lib, "nb200_set_pwm", I, vp, c_double_p, I, I, I, C.c_char_p)
    _sig(lib, "nb200_set_vmat", I, vp, c_double_p, I, I, I, I)
    _sig(lib, "nb200_set_fragment_sizes", I, vp, c_double_p, I)
    _sig(lib, "nb200_set_occ_model", I, vp, c_double_p, c_double_p, I, c_double_p, I, dbl)
    _sig(lib, "nb200_set_jitter", I, vp, c_double_p, i64)
    _sig(lib, "nb200_occ_configure", I, vp, C.POINTER(OccParams))
    _sig(lib, "nb200_nuc_configure", I, vp, C.POINTER(NucParams))
    _sig(lib, "nb200_fragmat_build", I, vp, c_int32_p, c_int32_p, i64, i32, i32, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_insertions", I, vp, c_int32_p, c_int32_p, i64, i32, i32, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_fragment_sizes", I, vp, i32, c_int32_p, c_int32_p, c_int64_p, c_int32_p, c_int32_p, i32, i32, i32,
         c_int64_p)
    _sig(lib, "nb200_bias_track", I, vp, c_uint8_p, i64, c_double_p)
    _sig(lib, "nb200_biasmat_build", I, vp, c_double_p, i64, i32, i32, c_double_p)
    _sig(lib, "nb200_get_ins", I, vp, c_double_p, i32, i32, i64, c_double_p)
    _sig(lib, "nb200_xcor_dense", I, vp, c_double_p, i64, c_double_p)
    _sig(lib, "nb200_coverage_dense", I, vp, c_double_p, i32, i64, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_smooth", I, vp, c_double_p, i64, c_double_p, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_call_peaks", I, vp, c_double_p, i64, dbl, i32, i32, i32, c_int32_p, i32, c_int32_p)
    _sig(lib, "nb200_reduce_peaks", I, vp, c_int32_p, c_double_p, i32, i32, c_int32_p)
    _sig(lib, "nb200_calculate_occupancy", I, vp, c_double_p, c_double_p, i32, c_double_p)
    _sig(lib, "nb200_multinomial_cov", I, vp, c_double_p, c_double_p, i64, i32, c_double_p)
    _sig(lib, "nb200_vplot", I, vp, i32, c_int32_p, c_int32_p, c_int64_p, c_int32_p, c_int32_p, i32, i32, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_coverage", I, vp, c_int32_p, c_int32_p, i64, i32, i32, i32, i32, i32, i32, c_double_p)
    _sig(lib, "nb200_batch_upload", I, vp, C.POINTER(Batch), C.POINTER(vp))
    _sig(lib, "nb200_batch_free", I, vp, vp)
    _sig(lib, "nb200_batch_sync", I, vp, vp)
    _sig(lib, "nb200_batch_total_len", i64, vp)
    _sig(lib, "nb200_batch_h2d_bytes", i64, vp)
    _sig(lib, "nb200_occ_run", I, vp, vp)
    _sig(lib, "nb200_nuc_run", I, vp, vp)
    _sig(lib, "nb200_occ_download", I, vp, vp, C.POINTER(OccOut))
    _sig(lib, "nb200_nuc_download", I, vp, vp, C.POINTER(NucOut))
    _sig(lib, "nb200_occ_d2h_bytes", i64, vp, C.POINTER(OccOut))
    _sig(lib, "nb200_nuc_d2h_bytes", i64, vp, C.POINTER(NucOut))
    _sig(lib, "nb200_occ_download32", I, vp, vp, C.POINTER(OccOut32))
    _sig(lib, "nb200_nuc_download32", I, vp, vp, C.POINTER(NucOut32))
    _sig(lib, "nb200_occ_d2h_bytes32", i64, vp, C.POINTER(OccOut32))
    _sig(lib, "nb200_nuc_d2h_bytes32", i64, vp, C.POINTER(NucOut32))
    _sig(lib, "nb200_timer_start", I, vp, vp)
    _sig(lib, "nb200_timer_stop", I, vp, vp)
    _sig(lib, "nb200_timer_elapsed_ms", I, vp, vp, C.POINTER(C.c_float))
    _sig(lib, "nb200_profile_enable", I, vp, I)
    _sig(lib, "nb200_profile_reset", I, vp)
    _sig(lib, "nb200_profile_count", I, vp)
    _sig(lib, "nb200_profile_get", I, vp, I, C.POINTER(C.c_char_p), c_int64_p, c_double_p)
    _sig(lib, "nb200_flush_l2", I, vp, vp)
    _sig(lib, "nb200_format_track", i64, C.c_char_p, i64, c_double_p, i64, i32, C.c_char_p, i64)
    _sig(lib, "nb200_bgzip_tabix", I, C.c_char_p, C.c_char_p, I, C.c_char_p, I)
    _sig(lib, "nb200_bgzip_tabix_level", I, C.c_char_p, C.c_char_p, I, I, C.c_char_p, I)
    _sig(lib, "nb200_bedgraph_fetch", I, C.c_char_p, C.c_char_p, i64, i64, dbl, c_double_p, C.c_char_p, I)
    _sig(lib, "nb200_tc_plan_describe", I, c_double_p, I, I, I, c_double_p, I, c_int32_p, c_int32_p, I)
    _sig(lib, "nb200_bam_fetch_many", I, C.c_char_p, i32, C.POINTER(C.c_uint64), c_int32_p, c_int32_p, c_int32_p, i32, c_int64_p,
         C.POINTER(c_int32_p), C.POINTER(c_int32_p), C.c_char_p, I)
    _sig(lib, "nb200_free", None, vp)
    _sig(lib, "nb200_nccl_unique_id", I, vp)
    _sig(lib, "nb200_nccl_init", I, vp, vp, I, I)
    _sig(lib, "nb200_allreduce_f64", I, vp, c_double_p, i64)
    _sig(lib, "nb200_allreduce_i64", I, vp, c_int64_p, i64)
    _sig(lib, "nb200_nccl_finalize", I, vp)
    _lib = lib
    return lib


EXPORTS = [
    "nb200_ctx_create", "nb200_ctx_destroy", "nb200_last_error", "nb200_device_info", "nb200_host_alloc",
    "nb200_host_free", "nb200_set_pwm", "nb200_set_vmat", "nb200_set_fragment_sizes", "nb200_set_occ_model",
    "nb200_set_jitter", "nb200_occ_configure", "nb200_nuc_configure", "nb200_fragmat_build", "nb200_insertions",
    "nb200_fragment_sizes", "nb200_bias_track", "nb200_biasmat_build", "nb200_get_ins", "nb200_xcor_dense",
    "nb200_coverage_dense", "nb200_smooth", "nb200_call_peaks", "nb200_reduce_peaks", "nb200_calculate_occupancy",
    "nb200_multinomial_cov", "nb200_batch_upload", "nb200_batch_free", "nb200_batch_sync", "nb200_batch_total_len",
    "nb200_batch_h2d_bytes", "nb200_occ_run", "nb200_nuc_run", "nb200_occ_download", "nb200_nuc_download",
    "nb200_occ_d2h_bytes", "nb200_nuc_d2h_bytes", "nb200_occ_download32", "nb200_nuc_download32", "nb200_occ_d2h_bytes32",
    "nb200_nuc_d2h_bytes32", "nb200_timer_start", "nb200_timer_stop", "nb200_timer_elapsed_ms",
    "nb200_profile_enable", "nb200_profile_reset", "nb200_profile_count", "nb200_profile_get", "nb200_flush_l2",
    "nb200_format_track", "nb200_bgzip_tabix", "nb200_bgzip_tabix_level", "nb200_bedgraph_fetch", "nb200_tc_plan_describe", "nb200_vplot", "nb200_coverage", "nb200_bam_fetch_many", "nb200_free",
    "nb200_nccl_unique_id", "nb200_nccl_init", "nb200_allreduce_f64", "nb200_allreduce_i64", "nb200_nccl_finalize",
]


def ptr(a, ctype):
    """numpy array -> typed ctypes pointer (None -> NULL).  The array must be C-contiguous."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    """int32 view/copy for the C-ABI; a value outside int32 raises instead of wrapping (numpy casts int64 arrays silently)."""
    if isinstance(a, np.ndarray) and a.dtype == np.int32:
        return np.ascontiguousarray(a)
    b = np.asarray(a)
    if b.size and b.dtype.kind in "iu" and b.dtype.itemsize > 4 and (int(b.min()) < -2 ** 31 or int(b.max()) >= 2 ** 31):
        raise OverflowError("coordinate outside int32: %d .. %d" % (int(b.min()), int(b.max())))
    return np.ascontiguousarray(b, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
