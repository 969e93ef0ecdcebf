"""ctypes binding of libdwb.so (include/dwb.h).  No torch types cross this boundary: only raw
device pointers, sizes and the CUDA stream handle.  There is no fallback of any kind: if the
library is missing or no sm_100 device is present, the first call raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DWB_LIB") or os.path.join(HERE, "libdwb.so")      # DWB_LIB: an alternative build of the same ABI (A/B runs)

DWB_MAX_POOL = 4
MODEL_WAVENET, MODEL_SASHIMI = 0, 1
F32, I64 = 0, 1


class DwbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdwb error {code}: {msg}")
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("unconditional", ctypes.c_int32),
        ("embed_in", ctypes.c_int32), ("embed_mid", ctypes.c_int32), ("embed_out", ctypes.c_int32),
        ("res_channels", ctypes.c_int32), ("skip_channels", ctypes.c_int32),
        ("num_res_layers", ctypes.c_int32), ("dilation_cycle", ctypes.c_int32),
        ("d_model", ctypes.c_int32), ("n_layers", ctypes.c_int32), ("n_pool", ctypes.c_int32),
        ("pool", ctypes.c_int32 * DWB_MAX_POOL), ("expand", ctypes.c_int32), ("ff", ctypes.c_int32),
        ("unet", ctypes.c_int32), ("L", ctypes.c_int32), ("d_state_half", ctypes.c_int32),
        ("mel_bands", ctypes.c_int32),
    ]


_P = ctypes.c_void_p
_I = ctypes.c_int
_I64 = ctypes.c_int64
_F = ctypes.c_float

# name -> argtypes, exactly the prototypes of include/dwb.h
SIGNATURES = {
    "dwb_version": [],
    "dwb_device_count": [ctypes.POINTER(_I)],
    "dwb_plan_create": [ctypes.POINTER(Config), _I, ctypes.POINTER(_P)],
    "dwb_plan_destroy": [_P],
    "dwb_plan_set_tensor": [_P, ctypes.c_char_p, _P, _I, ctypes.POINTER(_I64), _I, _I, _P],
    "dwb_plan_finalize": [_P, _P],
    "dwb_forward": [_P, _P, _P, _P, _I, _P, _I, _I, _P],
    "dwb_sample": [_P, _P, _P, _P, _I, ctypes.POINTER(_F), _I, _P, _I, _I, _I, _P],
    "dwb_sample_steps": [_P, _P, _P, _P, _I, ctypes.POINTER(_F), _I, _I, _I, _I, _I, _I, _P],
    "dwb_plan_cond_layout": [_P, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I64)],
    "dwb_plan_launch_count": [_P, ctypes.POINTER(_I64)],
    "dwb_plan_s4_blocks": [_P, ctypes.POINTER(_I)],
    "dwb_plan_s4_kernel": [_P, _I, _P, _I64, ctypes.POINTER(_I), ctypes.POINTER(_I)],
    "dwb_plan_mix_block": [_P, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "dwb_plan_work": [_P, _I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)],
    "dwb_plan_profile": [_P, _P, _P, _P, _I, _P, _I, _I, _I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64), _P],
    "dwb_cauchy_sym_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "dwb_cauchy_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "dwb_cauchy_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "dwb_cauchy_sym_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "dwb_s4_kernel_gen": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P],
    "dwb_fftconv_size": [_I, ctypes.POINTER(_I)],
    "dwb_fftconv_prepare": [_P, _P, _I, _I, _P, _P],
    "dwb_plan_cond_features": [_P, _P, _I, _I, _I, _P, _P],
    "dwb_debug_mix_trace": [_P, _I, _P, _P, _P, _P, _I, _P, _P],
    "dwb_debug_wave_trace": [_P, _I, _P, _P, _P, _P, _I, _I, _P, _P],
    "dwb_mel_frames": [_I, _I, _I, ctypes.POINTER(_I)],
    "dwb_mel_spectrogram": [_P, _I, _I, _F, _P, _I, _I, _P, _I, _F, _P, _P],
    "dwb_fftconv": [_P, _P, _P, _I64, _F, _F, _P, _P, _I, _I, _I, _P],
    "dwb_trainer_layout": [ctypes.POINTER(Config), _I, ctypes.c_char_p, _I, ctypes.POINTER(_I64), ctypes.POINTER(_I64),
                           ctypes.POINTER(_I), ctypes.POINTER(_I64)],
    "dwb_trainer_create": [ctypes.POINTER(Config), _I, _I, _I, ctypes.POINTER(_P)],
    "dwb_trainer_destroy": [_P],
    "dwb_trainer_info": [_P, ctypes.POINTER(_I64), ctypes.POINTER(_I64)],
    "dwb_trainer_set_gemm": [_P, _I],
    "dwb_trainer_loss_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "dwb_adam_step": [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I64, _F, _P],
}

PROF_CATEGORIES = ["embed", "init_conv", "head", "pool", "fftconv_s0", "fftconv_s1", "fftconv_s2", "fftconv_s3",
                   "mix_s0", "mix_s1", "mix_s2", "mix_s3", "wave_block"]

_lib = None


def lib():
    """Load libdwb.so (built in-tree by build.py / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m diffwave_sashimi_b200.build` "
                "(nvcc, sm_100a).  diffwave_sashimi_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _I
        L.dwb_last_error.argtypes = []
        L.dwb_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DwbError(rc, lib().dwb_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device (or host) address of a contiguous torch tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "libdwb needs contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
