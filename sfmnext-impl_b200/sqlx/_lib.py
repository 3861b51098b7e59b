"""ctypes binding of libsqlx.so (the C ABI declared in include/sqlx.h).

The product path has NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SQLX_LIB_PATH: development only -- load an A/B build of the same library (`make variant`, tools/ab.py)
LIB_PATH = os.environ.get("SQLX_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libsqlx.so")

c_float_p = ctypes.c_void_p   # device pointers are passed as raw addresses
c_size_t = ctypes.c_size_t
c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p

SQLX_AUTOMASK = 1
SQLX_AVG_REPROJ = 2
SQLX_NO_SSIM = 4
MAX_SOURCES = 4


class PhotoDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("S", ctypes.c_int32),
                ("ssim_radius", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("w_ssim", ctypes.c_float), ("w_l1", ctypes.c_float),
                ("noise_scale", ctypes.c_float), ("eps", ctypes.c_float)]


class ScaleDesc(ctypes.Structure):
    _fields_ = [("photo", PhotoDesc), ("Hc", ctypes.c_int32), ("Wc", ctypes.c_int32),
                ("smooth_weight", ctypes.c_float), ("rescale_translation", ctypes.c_int32)]


MAX_SCALES = 8


class MsDesc(ctypes.Structure):
    _fields_ = [("photo", PhotoDesc), ("num_scales", ctypes.c_int32),
                ("h", ctypes.c_int32 * MAX_SCALES), ("w", ctypes.c_int32 * MAX_SCALES),
                ("Hc", ctypes.c_int32 * MAX_SCALES), ("Wc", ctypes.c_int32 * MAX_SCALES),
                ("smooth_weight", ctypes.c_float * MAX_SCALES), ("rescale_translation", ctypes.c_int32)]


class PoseInputs(ctypes.Structure):
    _fields_ = [("axisangle", ctypes.c_void_p * MAX_SOURCES), ("translation", ctypes.c_void_p * MAX_SOURCES),
                ("fixed_T", ctypes.c_void_p * MAX_SOURCES), ("invert_mask", ctypes.c_uint32)]


_PROTOS = {
    "sqlx_last_error": (ctypes.c_char_p, []),
    "sqlx_version": (c_int, []),
    "sqlx_device_ok": (c_int, [c_int]),
    "sqlx_launch_count": (ctypes.c_ulonglong, []),
    "sqlx_profile_enable": (c_int, [c_int]),
    "sqlx_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
    "sqlx_depth_stats_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sqlx_depth_stats_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_depth_stats_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sqlx_reprojection_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p]),
    "sqlx_ssim_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_ssim_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sqlx_photo_workspace_bytes": (c_size_t, [ctypes.POINTER(PhotoDesc)]),
    "sqlx_photo_coef_bytes": (c_size_t, [ctypes.POINTER(PhotoDesc)]),
    "sqlx_pack_rgba": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_photo_fwd": (c_int, [ctypes.POINTER(PhotoDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_photo_bwd": (c_int, [ctypes.POINTER(PhotoDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "sqlx_photo_occ_workspace_bytes": (c_size_t, [ctypes.POINTER(PhotoDesc)]),
    "sqlx_photo_occ_fwd": (c_int, [ctypes.POINTER(PhotoDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                   ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_photo_occ_bwd": (c_int, [ctypes.POINTER(PhotoDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                   ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_float, c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "sqlx_warp_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sqlx_identity_losses_fwd": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_float,
                                         c_float, c_int, c_void_p, c_void_p]),
    "sqlx_scale_saved_bytes": (c_size_t, [ctypes.POINTER(ScaleDesc)]),
    "sqlx_scale_workspace_bytes": (c_size_t, [ctypes.POINTER(ScaleDesc)]),
    "sqlx_scale_loss_fwd": (c_int, [ctypes.POINTER(ScaleDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p,
                                    c_void_p, c_void_p, ctypes.POINTER(PoseInputs), c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "sqlx_scale_loss_bwd": (c_int, [ctypes.POINTER(ScaleDesc), c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p,
                                    c_void_p, c_void_p, ctypes.POINTER(PoseInputs), c_void_p, c_void_p, c_void_p, c_void_p,
                                    ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "sqlx_ms_saved_bytes": (c_size_t, [ctypes.POINTER(MsDesc)]),
    "sqlx_ms_workspace_bytes": (c_size_t, [ctypes.POINTER(MsDesc)]),
    "sqlx_ms_loss_fwd": (c_int, [ctypes.POINTER(MsDesc), ctypes.POINTER(c_void_p), c_void_p, ctypes.POINTER(c_void_p),
                                 ctypes.POINTER(c_void_p), c_void_p, c_void_p, ctypes.POINTER(PoseInputs), c_void_p,
                                 ctypes.POINTER(c_void_p), c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_size_t,
                                 c_void_p, c_size_t, c_void_p]),
    "sqlx_ms_loss_bwd": (c_int, [ctypes.POINTER(MsDesc), ctypes.POINTER(c_void_p), c_void_p, ctypes.POINTER(c_void_p),
                                 ctypes.POINTER(c_void_p), c_void_p, c_void_p, ctypes.POINTER(PoseInputs),
                                 ctypes.POINTER(c_void_p), c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                 ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "sqlx_rotation_warp_workspace_bytes": (c_size_t, [c_int]),
    "sqlx_rotation_warp_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_rotation_warp_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "sqlx_median_ratio_workspace_bytes": (c_size_t, [c_int]),
    "sqlx_median_ratio": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_median_ratio_resized": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                          c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_silog_workspace_bytes": (c_size_t, []),
    "sqlx_silog_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "sqlx_silog_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "sqlx_postprocess_disparity": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_head_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_head_linear_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "sqlx_head_centers_fwd": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "sqlx_head_centers_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "sqlx_backproject_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_backproject_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_project_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "sqlx_project_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "sqlx_smooth_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sqlx_smooth_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_smooth_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sqlx_pose_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_pose_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sqlx_sql_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "sqlx_sql_summary_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_summary_fwd_v1": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_tc_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "sqlx_sql_set_tensor_cores": (c_int, [c_int]),
    "sqlx_sql_get_tensor_cores": (c_int, []),
    "sqlx_sql_set_sm_budget": (c_int, [c_int]),
    "sqlx_sql_energy_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_sql_mix_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sqlx_sql_mix_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sqlx_sql_mix_weights_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "sqlx_sql_pred_mix_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p]),
    "sqlx_sql_bwd_pred_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_pred_mix_fwd_v1": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p]),
    "sqlx_sql_bwd_pred_mix_v1": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_bwd_summary": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_bwd_summary_v1": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_pred_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "sqlx_sql_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sqlx_sql_bwd_dx": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
}

_lib = None


class SqlxError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_PROTOS)


def lib():
    """Load libsqlx.so (once).  Raises SqlxError when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SqlxError("libsqlx.so not found at %s -- build it with `make` (or __graft_entry__.build()); "
                        "the sqlx hot path has no CPU / PyTorch fallback" % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in _PROTOS.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:                         # header / library mismatch: refuse to run on a partial library
        raise SqlxError("libsqlx.so at %s lacks symbols declared in include/sqlx.h: %s" % (LIB_PATH, ", ".join(missing)))
    _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().sqlx_last_error()
        raise SqlxError("%s failed (%d): %s" % (what, code, msg.decode() if msg else "?"))


def profile_enable(on=True):
    """Bracket every main kernel launch with CUDA events (bench.py roofline leg)."""
    check(lib().sqlx_profile_enable(int(bool(on))), "sqlx_profile_enable")


def profile_report():
    """{kernel name: (launches, total_ms)} for the launches recorded since the last report."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().sqlx_profile_report(buf, len(buf))
    if n < 0:
        check(n, "sqlx_profile_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    import torch
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise SqlxError("sqlx ops run on CUDA tensors only (got a %s tensor); there is no CPU path" % t.device)
        if t.dtype not in (torch.float32, torch.uint8):
            raise SqlxError("sqlx ops are fp32 (got %s)" % t.dtype)
