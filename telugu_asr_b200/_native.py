"""ctypes binding of include/tasr.h (libtasr_b200.so).  There is no other backend: if the
library cannot be loaded every operator in this package raises."""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtasr_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "tasr.h")

TASR_OK, TASR_ERR_BAD_ARG, TASR_ERR_UNSUPPORTED, TASR_ERR_MISALIGNED, TASR_ERR_CUDA = range(5)
ACT_CODES = {None: 0, "linear": 0, "none": 0, "tanh": 1, "gelu": 2, "relu": 3}
MATH_FP32, MATH_TF32 = 0, 1


class TasrFeatParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32), ("frame_length", C.c_int32), ("frame_step", C.c_int32),
        ("fft_length", C.c_int32), ("num_mel_bins", C.c_int32), ("normalize_signal", C.c_int32),
        ("log_base_e", C.c_int32), ("pad_end", C.c_int32), ("preemphasis", C.c_float),
        ("output_floor", C.c_float), ("feature_type", C.c_int32), ("normalize_zscore", C.c_int32),
        ("normalize_min_max", C.c_int32),
    ]


FEATURE_TYPES = {"log_mel_spectrogram": 0, "spectrogram": 1, "mfcc": 2, "waveform": 3}


class TasrDeferredGain(C.Structure):
    _fields_ = [("peak", C.c_void_p), ("log_scale_x2", C.c_float), ("log_floor", C.c_float)]


class TasrSepConvLayer(C.Structure):
    _fields_ = [
        ("dw", C.c_void_p), ("pw", C.c_void_p), ("bias", C.c_void_p),
        ("c_in", C.c_int32), ("c_out", C.c_int32), ("kernel", C.c_int32), ("stride", C.c_int32),
        ("same", C.c_int32), ("activation", C.c_int32),
    ]


class TasrEncoderBlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wq", "wk", "wv", "wo", "ln1_gamma", "ln1_beta", "w1", "b1", "w2", "b2", "ln2_gamma", "ln2_beta")] + [
        ("d_model", C.c_int32), ("num_heads", C.c_int32), ("head_dim", C.c_int32), ("fc_factor", C.c_int32), ("ln_eps", C.c_float)]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_SIGNATURES = {
    "tasr_version": (C.c_int, []),
    "tasr_last_error": (C.c_char_p, []),
    "tasr_launch_count": (C.c_int64, []),
    "tasr_featurizer_create": (C.c_int, [C.POINTER(TasrFeatParams), _vp, _vp, C.POINTER(_vp)]),
    "tasr_featurizer_destroy": (C.c_int, [_vp]),
    "tasr_featurizer_uses_fixed_mel": (C.c_int, [_vp]),
    "tasr_unpack_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "tasr_unpack_pcm16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "tasr_pack_valid_rows": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tasr_waveform_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "tasr_absmax_f32": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "tasr_logmel_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _vp]),
    "tasr_logmel_f32_lean": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _i32, _vp]),
    "tasr_logmel_f32_single_pass": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _i32, _vp,
                                              C.POINTER(TasrDeferredGain), _vp]),
    "tasr_apply_deferred_gain": (C.c_int, [_vp, _vp, _i32, _i32, _i32, C.POINTER(TasrDeferredGain), _vp]),
    "tasr_sepconv1d_f32": (C.c_int, [_vp, _i32, _i32, C.POINTER(TasrSepConvLayer), _vp, _i32, _vp]),
    "tasr_sepconv_plan_create": (C.c_int, [C.POINTER(TasrSepConvLayer), C.POINTER(_vp), _vp]),
    "tasr_sepconv_plan_destroy": (C.c_int, [_vp]),
    "tasr_sepconv1d_tf32": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "tasr_sepconv_plan_set_pad_row": (C.c_int, [_vp, _vp, _vp]),
    "tasr_sepconv_plan_pad_row": (C.c_void_p, [_vp]),
    "tasr_sepconv1d_tf32_ragged": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "tasr_sepconv_ragged_margin": (C.c_int32, []),
    "tasr_sepconv1d_tf32_ragged_lean": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32,
                                                  C.POINTER(TasrDeferredGain), _vp]),
    "tasr_conv_lengths_mask": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32),
                                         C.POINTER(_i32), _vp, _vp, _i32, _vp]),
    "tasr_conv2d_plan_create": (C.c_int, [_vp, _vp, _vp, _vp, _i32, C.POINTER(_vp), _vp]),
    "tasr_conv2d_plan_destroy": (C.c_int, [_vp]),
    "tasr_conv2d_output_shape": (C.c_int, [_i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "tasr_conv2d_subsample": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tasr_conv2d_plan_prepare_ragged": (C.c_int, [_vp, _i32, _vp]),
    "tasr_conv2d_subsample_ragged": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, C.POINTER(TasrDeferredGain), _vp]),
    "tasr_specaugment_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp]),
    "tasr_audio_mask": (C.c_int, [_vp, _i64, _i32, C.c_float, _vp, _vp]),
    "tasr_count_nonzero_frames": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "tasr_encoder_block_plan_create": (C.c_int, [C.POINTER(TasrEncoderBlockWeights), C.POINTER(_vp), _vp]),
    "tasr_encoder_block_plan_destroy": (C.c_int, [_vp]),
    "tasr_encoder_block_prepare": (C.c_int, [_vp, _i32]),
    "tasr_encoder_block_workspace_floats": (C.c_int64, [_vp, _i32, _i32]),
    "tasr_encoder_block_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def header_symbols() -> list[str]:
    """Every function include/tasr.h declares (used by the export test)."""
    with open(HEADER_PATH) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tasr_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    """Load libtasr_b200.so (building it in-tree with nvcc if it is absent).  Raises if neither
    is possible — there is no CPU or PyTorch fallback behind this package."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # always through build(): it returns at once when the source digest matches the stamp, and rebuilds a stale
        # library (a source, header or the generated mel_geometry.inc changed) instead of silently loading it
        from . import build as _build
        _build.build()
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError(
                f"telugu_asr_b200: cannot load {LIB_PATH}: {e}. Build it with "
                "`python -m telugu_asr_b200.build` (nvcc, sm_100a). There is no fallback path.") from e
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                raise RuntimeError(f"telugu_asr_b200: {LIB_PATH} does not export {name}; rebuild it")
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().tasr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Translate a TASR_ERR_* code into the exception the reference's layer would surface:
    bad shapes -> ValueError, unsupported configuration -> NotImplementedError, CUDA -> RuntimeError."""
    if rc == TASR_OK:
        return
    msg = last_error()
    if rc in (TASR_ERR_BAD_ARG, TASR_ERR_MISALIGNED):
        raise ValueError(msg)
    if rc == TASR_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def require_cuda(t, name: str):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(
            f"telugu_asr_b200: `{name}` must be a CUDA tensor — this package runs only on B200 "
            "(sm_100a) through libtasr_b200.so and has no CPU path")
    return t


poison_allocations = False


def empty(shape, dtype, device):
    """Output / intermediate allocation of the operators.  With `_native.poison_allocations = True` the buffer is
    filled with NaN (floats) or a large negative value (ints) first, so that a test can prove no kernel consumes a
    row that lean mode leaves unwritten."""
    import torch
    t = torch.empty(shape, dtype=dtype, device=device)
    if poison_allocations and t.numel():
        t.fill_(float("nan") if t.dtype.is_floating_point else -(2 ** 30))
    return t


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-stage CUDA-event marks (bench.py / tools): off unless `stage_marks` is a list ----------
stage_marks: list | None = None


def mark(name: str) -> None:
    """Record a CUDA event on the current stream under `name` when stage timing is switched on
    (`_native.stage_marks = []`); a no-op otherwise.  Consecutive marks bracket one stage."""
    if stage_marks is None:
        return
    import torch
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    stage_marks.append((name, ev))
