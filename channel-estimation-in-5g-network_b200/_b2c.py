"""ctypes binding of libb2c.so (include/b2c.h).  The product path has no CPU fallback: if the
library is missing or a tensor is not a contiguous CUDA tensor, calls raise."""

from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2c.so")

MAX_TAPS, MAX_ANT, MAX_SYM, N_OSC, N_STAT, N_BINSTAT = 16, 8, 16, 20, 3, 14
ABI_VERSION = 3
WIDE_PITCH = 600      # padded row pitch b2c_slot_pipeline accepts in its throughput configuration
EXPORTS = ["b2c_last_error_string", "b2c_abi_version", "b2c_tap_gains", "b2c_slot_pipeline",
           "b2c_ls_interp", "b2c_pilot_vectors", "b2c_mmse_dense", "b2c_dense_real_apply", "b2c_stats_bins",
           "b2c_ofdm_modulate", "b2c_ofdm_demodulate", "b2c_apply_channel", "b2c_tdl_full", "b2c_tdl_circular", "b2c_equalize",
           "b2c_qam_modulate", "b2c_qam_demodulate", "b2c_count_bit_errors", "b2c_bit_errors_per_slot", "b2c_pair00_moments", "b2c_pair00_errors", "b2c_count_nonfinite", "b2c_dense_prepared_bytes", "b2c_dense_prepare",
           "b2c_dense_apply_prepared", "b2c_dense_apply_grouped", "b2c_dense_score", "b2c_abs_diff_sum",
           "b2c_ml_features"]


class B2CError(RuntimeError):
    pass


class Geom(C.Structure):
    _fields_ = [("nsym", C.c_int32), ("nsc", C.c_int32), ("ntx", C.c_int32), ("nrx", C.c_int32),
                ("fft_size", C.c_int32), ("cp_length", C.c_int32), ("symbol_period_s", C.c_float),
                ("pitch", C.c_int32)]


class Profiles(C.Structure):
    _fields_ = [("n_models", C.c_int32), ("ntaps", C.c_void_p), ("npaths", C.c_void_p),
                ("tap_path", C.c_void_p), ("tap_amp", C.c_void_p), ("tap_tw", C.c_void_p),
                ("tap_corr", C.c_void_p), ("tap_delay", C.c_void_p)]


class Patterns(C.Structure):
    _fields_ = [("n_patterns", C.c_int32), ("np_max", C.c_int32), ("npilots", C.c_void_p),
                ("pilot_re", C.c_void_p), ("plan", C.c_void_p)]


class Slots(C.Structure):
    _fields_ = [("slot0", C.c_int64), ("seed", C.c_uint64), ("model_id", C.c_void_p),
                ("doppler_hz", C.c_void_p), ("snr_db", C.c_void_p), ("pattern_id", C.c_void_p), ("qpsk", C.c_int32)]


class DenseGroup(C.Structure):
    _fields_ = [("prepared", C.c_void_p), ("col0", C.c_int64), ("ncols", C.c_int64), ("np", C.c_int32)]


class PilotIO(C.Structure):
    _fields_ = [("hp", C.c_void_p), ("col", C.c_void_p), ("ld", C.c_int64)]


class Inject(C.Structure):
    _fields_ = [("jakes_u", C.c_void_p), ("p_max", C.c_int32), ("sym_turns", C.c_void_p),
                ("noise", C.c_void_p)]


_lib = None


def lib():
    """Load libb2c.so once.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2CError(f"{LIB_PATH} not found: build it with "
                           f"`python {os.path.join(HERE, 'build.py')}` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.b2c_last_error_string.restype = C.c_char_p
        L.b2c_abi_version.restype = C.c_int
        P, I64, I32, F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
        sig = {
            "b2c_tap_gains": [P, P, P, P, I64, P, P, P],
            "b2c_slot_pipeline": [P, P, P, P, P, I64, P, P, P, P, P, P, P, P, I32, P, P],
            "b2c_ls_interp": [P, P, P, P, I64, P, P, I64, P, I32, P, P, P, P, P, P, I64, P],
            "b2c_pilot_vectors": [P, P, I64, I32, F, I32, P, P],
            "b2c_mmse_dense": [P, I32, P, P, I64, I64, P],
            "b2c_dense_real_apply": [P, I32, I32, P, P, I64, I64, I64, P],
            "b2c_stats_bins": [P, P, P, P, I64, I32, P, P],
            "b2c_ofdm_modulate": [P, P, P, I64, P],
            "b2c_ofdm_demodulate": [P, P, P, I64, P],
            "b2c_apply_channel": [P, P, P, I64, P, P, P, P, P],
            "b2c_tdl_full": [P, P, I32, F, F, I64, I32, P, I32, P, C.c_uint64, I64, P, P],
            "b2c_tdl_circular": [P, P, P, I64, P, P, P, P],
            "b2c_equalize": [P, I64, P, P, P, C.c_double, I32, P],
            "b2c_qam_modulate": [P, I64, I32, P, P],
            "b2c_qam_demodulate": [P, I64, I32, I32, P, P],
            "b2c_count_bit_errors": [P, P, I64, P, P],
            "b2c_bit_errors_per_slot": [P, P, P, I64, P, P, I32, P, P],
            "b2c_pair00_moments": [P, I64, P, P, P, I64, P, P],
            "b2c_pair00_errors": [P, I64, P, P, I64, P, P, P],
            "b2c_count_nonfinite": [P, I64, I32, P, P],
            "b2c_abs_diff_sum": [P, P, I64, P, P],
            "b2c_dense_prepare": [P, I32, I32, I32, P, P],
            "b2c_dense_apply_prepared": [P, I32, I32, I32, P, P, I64, I64, I64, P],
            "b2c_dense_apply_grouped": [P, I32, P, P, I64, P],
            "b2c_dense_score": [P, P, P, P, I64, P, P, P, I64, P, P],
            "b2c_ml_features": [P, P, P, I64, P, P, P, I64, I32, I32, P, P, P, P],
        }
        L.b2c_dense_prepared_bytes.argtypes = [I32, I32, I32]
        L.b2c_dense_prepared_bytes.restype = C.c_int64
        for name, argtypes in sig.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        if L.b2c_abi_version() != ABI_VERSION:
            raise B2CError(f"libb2c ABI {L.b2c_abi_version()} != {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise B2CError(f"{what} failed ({rc}): {lib().b2c_last_error_string().decode()}")


_DT = {"c64": torch.complex64, "f32": torch.float32, "f64": torch.float64, "i32": torch.int32, "u8": torch.uint8,
       "c128": torch.complex128, "i64": torch.int64}


def dptr(t, kind, optional=False):
    """Device pointer of a contiguous CUDA tensor of the expected dtype."""
    if t is None:
        if optional:
            return None
        raise B2CError("required tensor is None")
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise B2CError("libb2c needs CUDA tensors (there is no CPU path)")
    if t.dtype != _DT[kind]:
        raise B2CError(f"expected dtype {_DT[kind]}, got {t.dtype}")
    if not t.is_contiguous():
        raise B2CError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def rows_ptr(t, pitch, optional=False):
    """Device pointer of a complex64 CUDA tensor whose rows (last dim) are `pitch` elements apart and whose
    leading dims are dense over those rows: a contiguous tensor (pitch = shape[-1]) or a [..., :nsc] view of one."""
    if t is None:
        if optional:
            return None
        raise B2CError("required tensor is None")
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex64):
        raise B2CError("libb2c needs complex64 CUDA tensors (there is no CPU path)")
    if t.is_contiguous() and pitch == t.shape[-1]:
        return C.c_void_p(t.data_ptr())
    st, sh = t.stride(), t.shape
    ok = st[-1] == 1 and (t.dim() < 2 or st[-2] == pitch) and sh[-1] <= pitch
    for i in range(t.dim() - 2):
        ok = ok and st[i] == st[i + 1] * sh[i + 1]
    if not ok:
        raise B2CError(f"tensor with shape {tuple(sh)} / strides {st} is not a dense stack of rows of pitch {pitch}")
    return C.c_void_p(t.data_ptr())


def row_pitch(t):
    """Row pitch (elements) of a dense stack of rows; shape[-1] for contiguous tensors."""
    return int(t.shape[-1]) if (t.is_contiguous() or t.dim() < 2) else int(t.stride(-2))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ref(struct):
    return C.cast(C.pointer(struct), C.c_void_p) if struct is not None else None
