"""ctypes binding of libcfr_b200.so (include/cfr_b200.h).  There is no CPU fallback: if the library or a
CUDA device is missing, using the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFR_LIB_PATH") or os.path.join(HERE, "libcfr_b200.so")   # env: A/B builds when profiling

MAX_PHASES, MAX_TAPS = 4, 9
ACT_NONE, ACT_LRELU, ACT_PRELU, ACT_RELU_POST = 0, 1, 2, 3

_i8_taps = (C.c_int8 * MAX_TAPS) * MAX_PHASES


class ConvDesc(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("N", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Cin", C.c_int32),
        ("w", C.c_void_p), ("wRows", C.c_int32), ("Kpad", C.c_int32),
        ("Cout", C.c_int32),
        ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("TW", C.c_int32), ("TH", C.c_int32), ("TN", C.c_int32),
        ("stride", C.c_int32), ("ntaps", C.c_int32), ("numPhases", C.c_int32),
        ("tap_dy", _i8_taps), ("tap_dx", _i8_taps),
        ("wRowsPerSample", C.c_int32), ("wRowsPerPhase", C.c_int32),
        ("out", C.c_void_p), ("outIsF32", C.c_int32), ("outH", C.c_int32), ("outW", C.c_int32),
        ("outC", C.c_int32), ("oscale", C.c_int32),
        ("ooff_y", C.c_int8 * MAX_PHASES), ("ooff_x", C.c_int8 * MAX_PHASES),
        ("bias", C.c_void_p),
        ("cbias", C.c_void_p), ("cbiasPerSample", C.c_int32),
        ("noise", C.c_void_p), ("noise_w", C.c_void_p),
        ("act", C.c_int32), ("slope", C.c_float), ("alpha", C.c_void_p),
        ("resid", C.c_void_p), ("residC", C.c_int32),
        ("stat_sum", C.c_void_p), ("stat_sq", C.c_void_p),
        ("kSplit", C.c_int32),
        ("keepMap", C.c_void_p), ("keepDim", C.c_int32),
    ]


class SamplerDesc(C.Structure):
    _fields_ = [
        ("synth", C.c_void_p), ("frm", C.c_void_p), ("chunk", C.c_int32),
        ("wp2", C.c_void_p), ("emb", C.c_void_p), ("dir_mat", C.c_void_p), ("w_avg", C.c_void_p),
        ("psi", C.c_float), ("gallery", C.c_void_p), ("n_gallery", C.c_int32),
        ("frm_group", C.c_int32), ("frm_big", C.c_void_p), ("emb_big", C.c_void_p), ("out_slot", C.c_void_p),
        ("matcher", C.c_void_p), ("tail", C.c_void_p),
        ("img_src", C.c_void_p), ("img_frm", C.c_void_p), ("img_chunk_bytes", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/cfr_b200.h declares
_P, _I, _F, _U64, _I64, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_int64, C.c_size_t
SIGNATURES = {
    "cfr_last_error": (C.c_char_p, []),
    "cfr_version": (_I, []),
    "cfr_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "cfr_program_create": (_I, [C.POINTER(_P)]),
    "cfr_program_destroy": (None, [_P]),
    "cfr_program_run": (_I, [_P, _P]),
    "cfr_program_num_launches": (_I, [_P]),
    "cfr_program_run_range": (_I, [_P, _I, _I, _P]),
    "cfr_program_op_label": (C.c_char_p, [_P, _I]),
    "cfr_program_op_flops": (C.c_double, [_P, _I]),
    "cfr_program_run_timed": (_I, [_P, _P, C.POINTER(C.c_float), _I]),
    "cfr_program_add_conv": (_I, [_P, C.POINTER(ConvDesc)]),
    "cfr_program_add_conv_halo": (_I, [_P, C.POINTER(ConvDesc), _P, _P]),
    "cfr_program_add_conv_halo_folded": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "cfr_program_add_upconv_blur_folded": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "cfr_program_add_memset": (_I, [_P, _P, _I, _SZ]),
    "cfr_program_add_styles": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "cfr_program_add_layer0": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "cfr_program_add_layer0_split": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "cfr_program_add_blur_act_stats_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I]),
    "cfr_program_add_affine_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _I]),
    "cfr_program_add_blur_act_stats": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I]),
    "cfr_program_add_finalize_stats": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P]),
    "cfr_program_add_affine": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "cfr_program_add_torgb_resize": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _F, _F, _P, _P, _P]),
    "cfr_program_add_torgb_resize_sparse": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _F, _F, _P, _P, _P, _P, _I]),
    "cfr_program_add_sum_partials": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "cfr_program_add_maxpool3s2": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _I]),
    "cfr_program_add_avgpool": (_I, [_P, _P, _I, _I, _I, _P]),
    "cfr_program_add_l2norm": (_I, [_P, _P, _I, _I, _P]),
    "cfr_noise_project": (_I, [_P, _P, _P, _I, _P, _P, _P, _F, _U64, _U64, _I, _P, _P, _P]),
    "cfr_truncate": (_I, [_P, _P, _F, _I, _P, _P]),
    "cfr_mapping": (_I, [_P, _P, _P, _I, _P, _P]),
    "cfr_match_vote": (_I, [_P, _I, _P, _I, _P, _P, _P, _P]),
    "cfr_matcher_create": (_I, [_P, _I, _I, _P, C.POINTER(_P)]),
    "cfr_matcher_destroy": (None, [_P]),
    "cfr_matcher_run": (_I, [_P, _P, _I, _P, _P, _P]),
    "cfr_match_keys": (_I, [_P, _I, _P, _I, C.c_uint32, _P, _P]),
    "cfr_matcher_keys": (_I, [_P, _P, _I, C.c_uint32, _P, _P]),
    "cfr_vote_keys": (_I, [_P, _I, _P, _P, _P]),
    "cfr_sampler_create": (_I, [C.POINTER(SamplerDesc), C.POINTER(_P)]),
    "cfr_sampler_set_overlap": (_I, [_P, _I]),
    "cfr_sampler_destroy": (None, [_P]),
    "cfr_sample_votes": (_I, [_P, _P, _P, _P, _I, _P, _I64, _U64, _U64, _P, _P, _P, _P, _P]),
    "cfr_sample_votes_multi": (_I, [_P, _I, _P, _P, _P, _I, _P, _U64, _P, _P, _P]),
    "cfr_sample_votes_host": (_I, [_P, _P, _P, _P, _I, _I64, _U64, _U64, _P, _P]),
    "cfr_launch_count": (_U64, []),
    "cfr_profile_enable": (_I, [_I]),
    "cfr_profile_read": (_I, [_I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (building is `python -m certifyingfacerecognition_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m certifyingfacerecognition_b200.build` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class CfrError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise CfrError(f"libcfr_b200 error {rc}: {load().cfr_last_error().decode()}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
