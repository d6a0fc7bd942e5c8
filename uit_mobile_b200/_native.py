"""ctypes binding of libuitk.so — the only way the Python host layer reaches the GPU kernels.

There is NO fallback: if the library is missing or a call fails, a ``UitkError`` is raised.  The signatures mirror
``include/uitk.h`` one to one.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libuitk.so")

PREC_FP32, PREC_BF16 = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}


class UitkError(RuntimeError):
    pass


ATTENTION = {"BNeckAttention": 0, "Attention": 1}
ACT = {"relu": 0, "gelu": 1}
POOLING = {"mean": 0, "token": 1, "dm": 2}


class EncoderCfg(C.Structure):
    """uitk_encoder_cfg (include/uitk.h)."""
    _fields_ = [("depth", C.c_int), ("outputdim", C.c_int), ("grid_t", C.c_int), ("precision", C.c_int),
                ("attention", C.c_int), ("act", C.c_int), ("pooling", C.c_int), ("reserved", C.c_int)]

    def __init__(self, depth=0, outputdim=0, grid_t=0, precision=0, attention=0, act=0, pooling=0):
        super().__init__(depth, outputdim, grid_t, precision, attention, act, pooling, 0)

    @property
    def tensor_core(self) -> bool:
        """The configuration the tcgen05 megakernel implements (UiT-XS/XXS/XXXS); everything else runs the fp32 kernels."""
        return self.precision == PREC_BF16 and self.attention == 0 and self.act == 0 and self.pooling == 0


# name -> (restype, argtypes); every symbol declared in include/uitk.h
SIGNATURES = {
    "uitk_version": (C.c_int, []),
    "uitk_last_error": (C.c_char_p, []),
    "uitk_kernel_launches": (C.c_uint64, []),
    "uitk_num_frames": (C.c_int64, [C.c_int64]),
    "uitk_num_crops": (C.c_int, [C.c_int64, C.c_int]),
    "uitk_tokens_per_crop": (C.c_int, [C.c_int64, C.c_int]),
    "uitk_tokens_total": (C.c_int, [C.POINTER(EncoderCfg), C.c_int64, C.c_int]),
    "uitk_frontend_blob_bytes": (C.c_size_t, [C.c_void_p]),
    "uitk_pack_frontend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "uitk_logmel": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uitk_logmel_i16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uitk_logmel_sliding_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "uitk_logmel_sliding": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "uitk_clamp_db": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "uitk_encoder_num_tensors": (C.c_int, [C.c_int]),
    "uitk_encoder_tensor_name": (C.c_char_p, [C.c_int, C.c_int]),
    "uitk_encoder_blob_bytes": (C.c_size_t, [C.POINTER(EncoderCfg)]),
    "uitk_pack_encoder": (C.c_int, [C.POINTER(EncoderCfg), C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t]),
    "uitk_encoder_workspace_bytes": (C.c_size_t, [C.POINTER(EncoderCfg), C.c_int64, C.c_int64, C.c_int]),
    "uitk_encoder": (C.c_int, [C.POINTER(EncoderCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uitk_encoder_fixup": (C.c_int, [C.POINTER(EncoderCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uitk_peer_words_slot_bytes": (C.c_size_t, []),
    "uitk_peer_words_publish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "uitk_peer_words_collect": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]),
    "uitk_init_bn": (C.c_int, [C.POINTER(EncoderCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "uitk_forward_features_workspace_bytes": (C.c_size_t, [C.POINTER(EncoderCfg), C.c_int64, C.c_int64]),
    "uitk_forward_features": (C.c_int, [C.POINTER(EncoderCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "uitk_forward_head": (C.c_int, [C.POINTER(EncoderCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "uitk_mnv2_num_tensors": (C.c_int, []),
    "uitk_mnv2_tensor_name": (C.c_char_p, [C.c_int]),
    "uitk_mnv2_blob_bytes": (C.c_size_t, [C.c_int]),
    "uitk_pack_mnv2": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t]),
    "uitk_mnv2_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "uitk_mnv2_forward": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uitk_encoder_tokens_offset": (C.c_size_t, [C.POINTER(EncoderCfg), C.c_int64, C.c_int64, C.c_int]),
    "uitk_debug_taps": (None, [C.c_int]),
    "uitk_debug_read_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "uitk_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "uitk_selftest_umma_ts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libuitk.so (built in-tree by ``python -m uit_mobile_b200.build``).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UitkError(f"{LIB_PATH} not found: build it with `python -m uit_mobile_b200.build` "
                            "(there is no CPU or PyTorch fallback for the UiT hot path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().uitk_last_error()
        raise UitkError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def encoder_tensor_names(depth: int):
    l = lib()
    return [l.uitk_encoder_tensor_name(depth, i).decode() for i in range(l.uitk_encoder_num_tensors(depth))]


def mnv2_tensor_names():
    l = lib()
    return [l.uitk_mnv2_tensor_name(i).decode() for i in range(l.uitk_mnv2_num_tensors())]
