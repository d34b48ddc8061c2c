"""ctypes binding of ``libdpde_b200.so`` (the C ABI declared in ``include/dpde_b200.h``).

PyTorch appears here only as the owner of device memory and streams: every call passes raw
``data_ptr()`` values and ``torch.cuda.current_stream().cuda_stream``.  There is no fallback:
if the library is missing the import of any op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPDE_B200_LIB") or os.path.join(_PKG, "lib", "libdpde_b200.so")   # the override is for kernel experiments

F32, F64, U8 = 0, 1, 2
PDE_NONE, PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL = 0, 1, 2, 3
NUM_SCALARS = 8

# every symbol include/dpde_b200.h declares (tests check the library exports exactly these)
EXPORTED = (
    "dpde_abi_version", "dpde_last_error", "dpde_guidance_workspace_bytes", "dpde_guidance_reduce",
    "dpde_guidance_finalize", "dpde_guidance_vjp", "dpde_laplacian", "dpde_sampler_init", "dpde_euler_predict",
    "dpde_euler_predict_bwd", "dpde_heun_guided_update", "dpde_heun_guided_update_rows", "dpde_halo_pack", "dpde_halo_unpack",
    "dpde_set_fast_path", "dpde_peer_alloc", "dpde_peer_free", "dpde_peer_export", "dpde_peer_open", "dpde_peer_close",
    "dpde_halo_push", "dpde_flag_wait", "dpde_set_tuning", "dpde_heat_residual_sq_workspace_bytes", "dpde_heat_residual_sq",
    "dpde_heat_residual_sq_vjp", "dpde_guidance_reduce_post", "dpde_mailbox_wait_finalize", "dpde_heun_guided_update_rows_push",
)
ABI_VERSION = 2
MAX_RANKS, MAILBOX_BYTES = 8, 512


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32), ("stride_b", C.c_int64),
                ("stride_c", C.c_int64)]


class GuidanceDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("ch_a", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("pde_kind", C.c_int32), ("has_a", C.c_int32), ("has_u", C.c_int32),
        ("slab_halo", C.c_int32), ("slab_row0", C.c_int32), ("slab_H_global", C.c_int32), ("_pad", C.c_int32),
        ("x0", View), ("dxdt", View), ("obs_a", View), ("mask_a", View), ("obs_u", View), ("mask_u", View),
        ("sample_coef", C.c_void_p), ("dx", C.c_double), ("w_a", C.c_double), ("w_u", C.c_double), ("w_pde", C.c_double),
        ("gamma", C.c_double), ("alpha", C.c_double), ("c_ex", C.c_double), ("c_an", C.c_double), ("tau", C.c_double),
        ("easy_axis", C.c_double * 3),
    ]


class Mailbox(C.Structure):
    """dpde_mailbox: cross-rank exchange of the three partial sums (boxes[r] = rank r's mailbox as mapped here)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_uint64), ("boxes", C.c_void_p * MAX_RANKS)]


class HaloPeers(C.Structure):
    """dpde_halo_peers: the neighbours' next-state buffers and flag words for the fused update + halo push."""
    _fields_ = [("up64", C.c_void_p), ("up32", C.c_void_p), ("down64", C.c_void_p), ("down32", C.c_void_p),
                ("flag_up", C.c_void_p), ("flag_down", C.c_void_p), ("ticket", C.c_void_p), ("epoch", C.c_uint64),
                ("H_up", C.c_int32), ("H_down", C.c_int32)]


class DpdeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DpdeError(f"{LIB_PATH} is missing: build it with `python -m dynamical_pde_diffusion_b200._build` "
                        "(or __graft_entry__.build()); there is no CPU or PyTorch fallback for these ops")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.dpde_abi_version.restype = C.c_int
    L.dpde_last_error.restype = C.c_char_p
    L.dpde_guidance_workspace_bytes.restype = C.c_size_t
    L.dpde_guidance_reduce.argtypes = [C.POINTER(GuidanceDesc), vp, vp, C.c_int, vp, vp, vp]
    L.dpde_guidance_finalize.argtypes = [C.POINTER(GuidanceDesc), vp, vp, vp, vp]
    L.dpde_guidance_vjp.argtypes = [C.POINTER(GuidanceDesc), vp, vp, vp, vp, vp]
    L.dpde_laplacian.argtypes = [vp, vp, i32, i64, i32, i32, i64, dbl, i32, vp]
    L.dpde_sampler_init.argtypes = [vp, dbl, vp, vp, i64, vp]
    L.dpde_euler_predict.argtypes = [vp, vp, dbl, dbl, vp, i64, vp]
    L.dpde_euler_predict_bwd.argtypes = [vp, dbl, dbl, vp, i64, vp]
    L.dpde_heun_guided_update.argtypes = [vp, vp, vp, vp, vp, dbl, dbl, vp, vp, i64, vp]
    L.dpde_heun_guided_update_rows.argtypes = [vp, vp, vp, vp, vp, dbl, dbl, vp, vp, i64, i64, i64, i64, vp]
    L.dpde_peer_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.dpde_peer_free.argtypes = [vp]
    L.dpde_peer_export.argtypes = [vp, C.c_char_p]
    L.dpde_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.dpde_peer_close.argtypes = [vp]
    L.dpde_halo_push.argtypes = [vp, i32, i64, i32, i32, i32, vp, i32, vp, i32, vp, vp, C.c_uint64, vp, vp]
    L.dpde_flag_wait.argtypes = [C.POINTER(vp), i32, C.c_uint64, dbl, vp, vp]
    L.dpde_halo_pack.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp, vp]
    L.dpde_halo_unpack.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp, vp]
    L.dpde_set_fast_path.argtypes = [C.c_int]
    L.dpde_set_tuning.argtypes = [C.c_int, C.c_int]
    L.dpde_heat_residual_sq_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.dpde_heat_residual_sq_workspace_bytes.restype = C.c_size_t
    L.dpde_heat_residual_sq.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, vp, dbl, vp, vp, vp]
    L.dpde_heat_residual_sq_vjp.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, vp, dbl, vp, vp, vp, vp]
    L.dpde_guidance_reduce_post.argtypes = [C.POINTER(GuidanceDesc), vp, vp, C.POINTER(Mailbox), vp]
    L.dpde_mailbox_wait_finalize.argtypes = [C.POINTER(GuidanceDesc), C.POINTER(Mailbox), dbl, vp, vp, vp, vp, vp]
    L.dpde_heun_guided_update_rows_push.argtypes = [vp, vp, vp, vp, vp, dbl, dbl, vp, vp, i64, i32, i32, i32, C.POINTER(HaloPeers), vp]
    for name in EXPORTED:
        fn = getattr(L, name)
        if name not in ("dpde_abi_version", "dpde_last_error", "dpde_guidance_workspace_bytes",
                        "dpde_heat_residual_sq_workspace_bytes"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise DpdeError(f"dpde_b200 error {rc}: {lib().dpde_last_error().decode(errors='replace')}")


# number of kernels this process launched through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0
# when set to a list, every launch is bracketed by CUDA events on the current stream: (name, start, stop)
event_log = None


def call(name: str, *args) -> None:
    global launch_count
    if event_log is None:
        check(getattr(lib(), name)(*args))
    else:
        import torch

        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(getattr(lib(), name)(*args))
        b.record()
        event_log.append((name, a, b))
    launch_count += 1
