"""ctypes binding of libdm_b200.so -- one-to-one with include/dm_abi.h."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdm_b200.so")

P = ctypes.c_void_p
I = ctypes.c_int
L = ctypes.c_int64
F = ctypes.c_float
D = ctypes.c_double
PI = ctypes.POINTER(ctypes.c_int)
PD = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); every symbol include/dm_abi.h declares
SIGNATURES = {
    "dm_last_error": (ctypes.c_char_p, []),
    "dm_abi_version": (I, []),
    "dm_create": (I, [I, ctypes.POINTER(P)]),
    "dm_destroy": (I, [P]),
    "dm_load_tensor": (I, [P, ctypes.c_char_p, P, I, I, ctypes.POINTER(L)]),
    "dm_finalize_weights": (I, [P]),
    "dm_save_packed": (I, [P, ctypes.c_char_p]),
    "dm_load_packed": (I, [P, ctypes.c_char_p]),
    "dm_set_schedule": (I, [P, P, P, I]),
    "dm_set_context": (I, [P, I, P, P]),
    "dm_vae_encode": (I, [P, P, P, I, I, I, P, P, P, P]),
    "dm_unet_eps": (I, [P, P, P, P, I, I, I, P, P]),
    "dm_unet_rows": (I, [P, P, P, P, P, P, P, I, I, I, P, P, I, P]),
    "dm_compute_loss": (I, [P, P, P, P, P, I, I, I, I, P, P]),
    "dm_typicality": (I, [P, P, P, P, P, I, I, I, I, I, P, P, I, P]),
    "dm_dift": (I, [P, P, P, L, I, I, I, I, I, I, P, P]),
    "dm_dift_shape": (I, [I, I, I, PI, PI, PI]),
    "dm_launch_count": (L, [P]),
    "dm_flop_count": (D, [P]),
    "dm_debug_keep": (I, [P, I]),
    "dm_debug_fetch": (L, [P, ctypes.c_char_p, P, L, PI, P]),
    "dm_profile_unet": (I, [P, I, I, I, I, PD, PD, PD, PD, PD]),
    "dm_profile_plan": (I, [P, I, I, I, I, I, I, PD, PD]),
    "dm_op_conv": (I, [P, P, I, I, I, I, I, P, I, I, I, I, P, P, P, P, I, I, I, I, P]),
    "dm_op_conv_gn": (I, [P, I, I, I, I, P, I, I, P, P, P, P, P, F, I, P, P, P]),
    "dm_op_attention": (I, [P, P, P, L, L, L, L, L, L, I, I, I, I, I, I, P, P, L, P]),
    "dm_op_groupnorm": (I, [P, P, I, I, I, I, P, P, F, I, P, P]),
    "dm_op_layernorm": (I, [P, L, I, P, P, F, P, P]),
    "dm_op_set_variant": (I, [ctypes.c_char_p, I]),
    "dm_patch_topk": (I, [P, I, I, I, I, I, I, I, I, I, P, P, P, P, P]),
    "dm_patch_topk_work_floats": (L, [I, I, I, I]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the engine library (built in-tree by __graft_entry__.build() / `make -C diff-mining_b200/csrc`).
    Fails loudly when it is missing -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()')")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("dm_b200: " + load().dm_last_error().decode(errors="replace"))
