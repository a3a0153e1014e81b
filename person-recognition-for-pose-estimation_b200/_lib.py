"""ctypes binding of libspp.so (the C ABI declared in include/spp.h).

The library is built in-tree by ``csrc/Makefile`` (``build()`` below, also called from
``__graft_entry__.build``).  There is deliberately no fallback: if the shared object is missing or
there is no CUDA device, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_size_t, c_ubyte, c_uint16, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspp.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_PEERS = 16            # SPP_MAX_PEERS
IPC_HANDLE_BYTES = 64     # SPP_IPC_HANDLE_BYTES


class PeerGroupStruct(Structure):
    """``spp_peer_group`` of include/spp.h."""
    _fields_ = [("world", c_int), ("rank", c_int), ("m_local", c_int), ("buffers", c_void_p * MAX_PEERS)]


# name -> (restype, argtypes); mirrors include/spp.h and include/spp_internal.h
_P = c_void_p
SIGNATURES = {
    "spp_abi_version": (c_int, []),
    "spp_last_error": (c_char_p, []),
    "spp_device_sm_count": (c_int, []),
    "spp_set_launch_limit": (c_int, [c_int, c_int]),
    "spp_head_decode": (c_int, [POINTER(_P), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, _P, _P]),
    "spp_head_decode_split": (c_int, [POINTER(_P), POINTER(_P), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int,
                                      _P, _P]),
    "spp_decode_nms_split": (c_int, [POINTER(_P), POINTER(_P), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int,
                                     c_int, c_float, c_float, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "spp_decode_nms_mode": (c_int, [c_int]),
    "spp_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "spp_nms_decoded": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_float, c_int, _P, _P, _P, _P,
                                c_size_t, _P]),
    "spp_decode_nms": (c_int, [POINTER(_P), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, c_float,
                               c_float, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "spp_l2_normalize": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "spp_f32_to_bf16": (c_int, [_P, c_size_t, _P, _P]),
    "spp_match_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spp_match_top1": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "spp_match_top1_ex": (c_int, [_P, _P, _P, c_float, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "spp_match_unpack_keys": (c_int, [_P, c_int, c_float, _P, _P, _P]),
    "spp_peer_buffer_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spp_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p), POINTER(c_ubyte)]),
    "spp_peer_open": (c_int, [POINTER(c_ubyte), POINTER(c_void_p)]),
    "spp_peer_close": (c_int, [_P]),
    "spp_peer_free": (c_int, [_P]),
    "spp_peer_can_access": (c_int, [c_int]),
    "spp_sharded_match_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "spp_sharded_match_top1": (c_int, [POINTER(PeerGroupStruct), _P, _P, _P, c_float, c_int, c_int, c_int, c_float, c_int, _P, _P, _P,
                                       _P, c_size_t, _P]),
    "spp_associate": (c_int, [_P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "spp_crop_affine": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                c_int, _P, _P]),
    "spp_crop_affine_u8": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                   c_int, _P, _P]),
    "spp_crop_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "spp_crop_policy": (c_int, [c_int]),
    "spp_crop_affine_ws": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                   c_int, _P, _P, c_size_t, _P]),
    "spp_crop_affine_ex": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), c_int,
                                   _P, c_int, _P, c_size_t, c_int, _P]),
    "spp_crop_plan": (c_int, [c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "spp_crop_affine_run": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                    c_int, _P, _P, c_size_t, _P]),
    "spp_crop_affine_u8_run": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                       c_int, _P, _P, c_size_t, _P]),
    "spp_crop_affine_u8_ws": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                      c_int, _P, _P, c_size_t, _P]),
    "spp_heatmap_decode": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P,
                                   _P, _P]),
    "spp_heatmap_decode_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P,
                                        _P, _P]),
    "spp_pose_results": (c_int, [_P, _P, _P, c_int, c_int, c_float, _P, _P, _P]),
    "spp_pose_oks": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, _P, _P]),
    "spp_det_match_targets": (c_int, [_P, _P, c_int, _P, _P, c_int, POINTER(c_float), c_int, c_int, _P, _P]),
    "spp_det_ap_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spp_det_average_precision": (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, c_int, c_double, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    # test hook (include/spp_internal.h)
    "spp_debug_match_top1_simt": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
}

_lib = None


class SppError(RuntimeError):
    """Raised when a libspp entry point returns an error code (message from spp_last_error)."""


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libspp.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libspp.so failed (see output above)")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SppError(f"{LIB_PATH} is missing: run `make -C {CSRC}` (or __graft_entry__.build()); "
                           "there is no CPU / PyTorch fallback for the spp ops")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().spp_last_error().decode("utf-8", "replace")
        raise SppError(f"{what} failed (code {rc}): {msg}")
