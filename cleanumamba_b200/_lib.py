"""ctypes binding of libcleanumamba_sm100.so (the C ABI declared in include/cleanumamba_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
PyTorch is used only as the owner of device memory and streams; raw device pointers cross the boundary.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUM_LIB_PATH") or os.path.join(_PKG, "libcleanumamba_sm100.so")    # override: A/B of two builds

# enums (mirror include/cleanumamba_b200.h)
EPI_NONE, EPI_RELU, EPI_SILU = 0, 1, 2
EPI_GLU = {"Sigmoid": 8, "ReLU": 9, "SiLU": 10, "GELU": 11}
MATH_FP32, MATH_TF32X3, MATH_TF32, MATH_BF16X3, MATH_F16X3, MATH_BF16 = 0, 1, 2, 3, 4, 5
MATH_BY_NAME = {"fp32": MATH_FP32, "tf32x3": MATH_TF32X3, "tf32": MATH_TF32, "bf16x3": MATH_BF16X3, "f16x3": MATH_F16X3,
                "bf16": MATH_BF16}


class GemmDesc(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a_batch_stride", C.c_longlong), ("a_row_stride", C.c_longlong),
                ("a_rows", C.c_int), ("k", C.c_int), ("taps", C.c_int), ("tap_shift", C.c_int * 2),
                ("w", C.c_void_p), ("ldw", C.c_int), ("bias", C.c_void_p),
                ("c", C.c_void_p), ("c_batch_stride", C.c_longlong), ("c_row_stride", C.c_longlong),
                ("m", C.c_int), ("n", C.c_int), ("batch", C.c_int), ("epilogue", C.c_int),
                ("addend", C.c_void_p), ("add_batch_stride", C.c_longlong), ("add_row_stride", C.c_longlong),
                ("math", C.c_int), ("w_lo", C.c_void_p), ("acc_scale", C.c_float), ("out_bf16", C.c_int), ("w_lo_is_zero", C.c_int),
                ("a_lo", C.c_void_p), ("c_lo", C.c_void_p), ("addend_lo", C.c_void_p), ("a_scale_dev", C.c_void_p), ("addend_is_mask", C.c_int), ("aux", C.c_void_p), ("aux_batch_stride", C.c_longlong), ("aux_row_stride", C.c_longlong),
                ("cta_pair", C.c_int),
                ("a_planes", C.c_int), ("a_plane_k", C.c_int), ("a_plane0", C.c_int), ("a_plane_step", C.c_int), ("n_half", C.c_int),
                ("small_m_path", C.c_int)]


class Enc0BlockDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_stride", C.c_longlong), ("batch", C.c_int), ("length", C.c_int),
                ("conv_w", C.c_void_p), ("conv_b", C.c_void_p), ("glu_w_hi", C.c_void_p), ("glu_w_lo", C.c_void_p),
                ("glu_b", C.c_void_p), ("acc_scale", C.c_float), ("w_lo_is_zero", C.c_int),
                ("out", C.c_void_p), ("rows_out", C.c_int), ("channels", C.c_int), ("channels_out", C.c_int)]


class DecLastBlockDesc(C.Structure):
    _fields_ = [("a", C.c_void_p), ("batch", C.c_int), ("rows_in", C.c_int),
                ("glu_w_hi", C.c_void_p), ("glu_w_lo", C.c_void_p), ("glu_b", C.c_void_p), ("acc_scale", C.c_float),
                ("w_lo_is_zero", C.c_int), ("convt_w", C.c_void_p), ("convt_bias", C.c_float), ("scale", C.c_void_p),
                ("out", C.c_void_p), ("out_stride", C.c_longlong), ("out_length", C.c_int), ("channels", C.c_int),
                ("channels_gated", C.c_int)]


class ScanDesc(C.Structure):
    _fields_ = [("u", C.c_void_p), ("u_bs", C.c_longlong), ("u_rs", C.c_longlong),
                ("delta", C.c_void_p), ("dl_bs", C.c_longlong), ("dl_rs", C.c_longlong),
                ("z", C.c_void_p), ("z_bs", C.c_longlong), ("z_rs", C.c_longlong),
                ("Bm", C.c_void_p), ("B_bs", C.c_longlong), ("B_rs", C.c_longlong),
                ("Cm", C.c_void_p), ("C_bs", C.c_longlong), ("C_rs", C.c_longlong),
                ("y", C.c_void_p), ("y_bs", C.c_longlong), ("y_rs", C.c_longlong),
                ("a2", C.c_void_p), ("Dskip", C.c_void_p), ("delta_bias", C.c_void_p),
                ("h0", C.c_void_p), ("h_out", C.c_void_p),
                ("batch", C.c_int), ("len", C.c_int), ("d", C.c_int), ("n_state", C.c_int),
                ("delta_softplus", C.c_int), ("h_ckpt", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
                ("state_f16", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("dz", C.c_void_p), ("dz_batch_stride", C.c_longlong), ("dz_row_stride", C.c_longlong),
                ("a", C.c_void_p), ("a_batch_stride", C.c_longlong), ("a_row_stride", C.c_longlong), ("a_rows", C.c_int),
                ("dw", C.c_void_p), ("ldw", C.c_int), ("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("taps", C.c_int),
                ("tap_shift", C.c_int * 2), ("batch", C.c_int), ("math", C.c_int), ("workspace", C.c_void_p), ("dz_scale_dev", C.c_void_p)]


class ScanBwdDesc(C.Structure):
    _fields_ = [("fwd", ScanDesc), ("h_ckpt", C.c_void_p),
                ("dout", C.c_void_p), ("dout_bs", C.c_longlong), ("dout_rs", C.c_longlong),
                ("du", C.c_void_p), ("du_bs", C.c_longlong), ("du_rs", C.c_longlong),
                ("ddelta", C.c_void_p), ("ddl_bs", C.c_longlong), ("ddl_rs", C.c_longlong),
                ("dz", C.c_void_p), ("dz_bs", C.c_longlong), ("dz_rs", C.c_longlong),
                ("dB", C.c_void_p), ("dB_bs", C.c_longlong), ("dB_rs", C.c_longlong),
                ("dC", C.c_void_p), ("dC_bs", C.c_longlong), ("dC_rs", C.c_longlong),
                ("dA_log", C.c_void_p), ("dD", C.c_void_p), ("ddelta_bias", C.c_void_p)]


class ShiftEntry(C.Structure):
    _fields_ = [("base", C.c_void_p), ("row_stride", C.c_longlong), ("src_off", C.c_longlong), ("count", C.c_longlong),
                ("rows", C.c_int), ("reserved", C.c_int)]


EXPORTS = {
    "cum_abi_version": (C.c_int, []),
    "cum_init": (C.c_int, [C.c_int]),
    "cum_shutdown": (C.c_int, []),
    "cum_last_error": (C.c_char_p, []),
    "cum_wave_normalize_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cum_conv_in_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cum_conv_in_bf16_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_conv_in_hl16_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_convt_out_bf16_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_stream_std_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "cum_stream_std_counter_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
    "cum_convt_out_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_conv_in_strided_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p]),
    "cum_convt_out_strided_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_dwconv_silu_strided_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p]),
    "cum_stream_shift_fwd": (C.c_int, [C.POINTER(ShiftEntry), C.c_int, C.c_void_p]),
    "cum_gemm_bias_act_fwd": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "cum_enc0_block_fwd": (C.c_int, [C.POINTER(Enc0BlockDesc), C.c_void_p]),
    "cum_dec_last_block_fwd": (C.c_int, [C.POINTER(DecLastBlockDesc), C.c_void_p]),
    "cum_split_tf32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cum_split_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cum_split_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]),
    "cum_ln_residual_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "cum_dwconv_silu_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_selective_scan_fwd": (C.c_int, [C.POINTER(ScanDesc), C.c_void_p]),
    "cum_selective_scan_workspace_bytes": (C.c_longlong, [C.POINTER(ScanDesc)]),
    "cum_glu_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "cum_glu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "cum_relu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "cum_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "cum_add_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cum_gemm_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "cum_grad_scale_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cum_gemm_wgrad_workspace_bytes": (C.c_longlong, [C.POINTER(WgradDesc)]),
    "cum_ln_residual_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "cum_dwconv_silu_bwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]),
    "cum_conv_in_bwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cum_convt_out_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cum_selective_scan_bwd": (C.c_int, [C.POINTER(ScanBwdDesc), C.c_void_p]),
    "cum_stft_frames_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "cum_stft_loss_reduce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cum_stft_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cum_stft_overlap_add": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong,
                                       C.c_void_p]),
    "cum_channel_importance_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited_devices = set()


def load() -> C.CDLL:
    """dlopen the library (no CUDA call is made here, so this also works on a GPU-less host)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m cleanumamba_b200.build` "
                        "(cleanumamba_b200 has no CPU / PyTorch fallback)")
                lib = C.CDLL(LIB_PATH)
                for name, (res, args) in EXPORTS.items():
                    fn = getattr(lib, name)
                    fn.restype, fn.argtypes = res, args
                _lib = lib
    return _lib


def last_error() -> str:
    return load().cum_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def init(device: torch.device) -> C.CDLL:
    lib = load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        check(lib.cum_init(idx), "cum_init")
        _inited_devices.add(idx)
    return lib


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
