"""ctypes binding of librtpose_b200.so (C ABI declared in include/rtpose_b200.h).

There is deliberately NO fallback: if the shared library is missing, or a compute call is made without an
sm_100 device, this raises.  (`load()` itself works on a CPU-only box so the symbol table can be checked.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTPOSE_B200_LIB") or os.path.join(HERE, "librtpose_b200.so")  # override: A/B runs of two builds
MAX_TAPS = 27
GUARD_BYTES = 8192


class P8Struct(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n_stride", C.c_int64), ("c_stride", C.c_int64), ("N", C.c_int32),
                ("C8", C.c_int32), ("Z", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32)]


NULL_P8 = P8Struct(None, 0, 0, 0, 0, 0, 0, 0)
RTP_LOSS_SPARSE_DREG = 1  # include/rtpose_b200.h


class ConvDesc(C.Structure):
    _fields_ = [("inp", P8Struct), ("out", P8Struct), ("res", P8Struct), ("mask", P8Struct), ("w", C.c_void_p),
                ("bias", C.c_void_p), ("Cin", C.c_int32), ("NP", C.c_int32), ("out_c8", C.c_int32),
                ("ntaps", C.c_int32), ("tz", C.c_int8 * MAX_TAPS), ("tx", C.c_int8 * MAX_TAPS),
                ("ty", C.c_int8 * MAX_TAPS), ("wt", C.c_int8 * MAX_TAPS), ("RZ", C.c_int32), ("RX", C.c_int32),
                ("RY", C.c_int32), ("IS", C.c_int32), ("OS", C.c_int32), ("oz0", C.c_int32), ("ox0", C.c_int32),
                ("oy0", C.c_int32), ("relu", C.c_int32), ("accumulate", C.c_int32)]


class ConvK3S1Desc(C.Structure):
    _fields_ = [("inp", P8Struct), ("out", P8Struct), ("res", P8Struct), ("mask", P8Struct), ("w", C.c_void_p),
                ("bias", C.c_void_p), ("Cin", C.c_int32), ("NPo", C.c_int32), ("out_c8", C.c_int32),
                ("relu", C.c_int32), ("accumulate", C.c_int32), ("stat_mode", C.c_int32), ("stat_aux", P8Struct),
                ("stat_ws", C.c_void_p), ("use_tap_mask", C.c_int32), ("tap_mask_groups", C.c_int32), ("tap_mask", C.c_uint16 * 8),
                ("debug", C.c_void_p), ("unit_list", C.c_void_p), ("unit_count", C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [("x", P8Struct), ("dy", P8Struct), ("Cin", C.c_int32), ("NP", C.c_int32), ("ntaps", C.c_int32),
                ("tz", C.c_int8 * MAX_TAPS), ("tx", C.c_int8 * MAX_TAPS), ("ty", C.c_int8 * MAX_TAPS),
                ("RZ", C.c_int32), ("RX", C.c_int32), ("RY", C.c_int32), ("IS", C.c_int32), ("nsplit", C.c_int32),
                ("tc", C.c_int16 * MAX_TAPS), ("workspace", C.c_void_p)]


class FuseDesc(C.Structure):
    _fields_ = [("out", P8Struct), ("C", C.c_int32), ("n_same", C.c_int32), ("n_low", C.c_int32),
                ("same", P8Struct * 4), ("low", P8Struct * 3), ("bias", C.c_void_p), ("relu", C.c_int32)]


class ConatDesc(C.Structure):
    _fields_ = [("x0", P8Struct), ("out", P8Struct), ("low", P8Struct * 3), ("n_low", C.c_int32), ("c_x0", C.c_int32),
                ("c_low", C.c_int32 * 3), ("w", C.c_void_p), ("bias", C.c_void_p), ("K", C.c_int32), ("NP", C.c_int32),
                ("out_c8", C.c_int32), ("relu", C.c_int32)]


class PackJob(C.Structure):  # rtp_pack_job
    _fields_ = [("w", C.c_void_p), ("dst", C.c_void_p), ("kind", C.c_int32), ("Cout", C.c_int32), ("Cin_total", C.c_int32),
                ("ntaps", C.c_int32), ("ci0", C.c_int32), ("ci_n", C.c_int32), ("KP", C.c_int32), ("NP", C.c_int32),
                ("flag", C.c_int32), ("block0", C.c_int32), ("nblocks", C.c_int32), ("reserved", C.c_int32)]


class NpyInfo(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("elem_bytes", C.c_int32), ("fortran_order", C.c_int32), ("reserved_", C.c_int32),
                ("shape", C.c_int64 * 8), ("data_offset", C.c_int64), ("file_bytes", C.c_int64), ("descr", C.c_char * 16)]


_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> (restype, argtypes); every symbol include/rtpose_b200.h declares
PROTOTYPES = {
    "rtp_last_error": (C.c_char_p, []),
    "rtp_version": (C.c_int, []),
    "rtp_device_ok": (C.c_int, []),
    "rtp_set_shared_carveout": (C.c_int, [_i32]),
    "rtp_npy_probe": (C.c_int, [C.c_char_p, C.POINTER(NpyInfo)]),
    "rtp_npy_roi_slab_bytes": (C.c_int64, [C.POINTER(NpyInfo), _i32, _i32]),
    "rtp_npy_read_roi_slab": (C.c_int, [C.c_char_p, _i32, _i32, _i32, _i32, _vp, _i64, _i32]),
    "rtp_pack_ncdhw": (C.c_int, [_vp, P8Struct, _i32, _vp]),
    "rtp_unpack_ncdhw": (C.c_int, [P8Struct, _vp, _i32, _i32, _vp]),
    "rtp_ingest_pack": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, P8Struct,
                                  _vp, _vp]),
    "rtp_weight_pack": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_conv": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "rtp_conv_multi": (C.c_int, [C.POINTER(ConvDesc), _i32, _vp]),
    "rtp_weight_pack_batch": (C.c_int, [_vp, _i32, _i32, _vp]),
    "rtp_weight_pack_k3s1": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_weight_pack_k3s1_window": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_conv_k3s1": (C.c_int, [C.POINTER(ConvK3S1Desc), _vp]),
    "rtp_conv_k3s1_smem_bytes": (C.c_int64, [_i32, _i32, _i32, _i32, _i32]),
    "rtp_conv_k3s1_stat_ws_bytes": (C.c_int64, [_i32]),
    "rtp_conv_k3s1_num_ctas": (C.c_int32, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "rtp_conv_k3s1_stat_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _f32, _vp, _vp, _vp]),
    "rtp_conv_pw_supported": (C.c_int, [_i32, _i32]),
    "rtp_conv_pw": (C.c_int, [P8Struct, P8Struct, P8Struct, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_wgrad_workspace_bytes": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "rtp_wgrad": (C.c_int, [C.POINTER(WgradDesc), _vp]),
    "rtp_wgrad_reduce": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_wgrad_k3s1_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "rtp_wgrad_k3s1_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_wgrad_k3s1_zero_bytes": (C.c_int64, [_i32]),
    "rtp_wgrad_k3s1": (C.c_int, [P8Struct, P8Struct, _i32, _vp, _vp, C.POINTER(_i32), _vp]),
    "rtp_wgrad_k3s1_units": (C.c_int, [P8Struct, P8Struct, _i32, _vp, _vp, C.POINTER(_i32), _vp, _vp, _vp]),
    "rtp_active_units": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "rtp_wgrad_k3s1_reduce": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rtp_gn_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_gn_sums": (C.c_int, [P8Struct, _i32, _vp, _vp, _vp]),
    "rtp_gn_stats": (C.c_int, [P8Struct, _i32, _i32, _f32, _vp, _vp, _vp]),
    "rtp_gn_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _f32, _vp, _vp]),
    "rtp_gn_apply": (C.c_int, [P8Struct, _i32, _i32, _vp, _vp, _vp, P8Struct, _vp]),
    "rtp_gn_bwd_reduce": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rtp_gn_bwd_apply": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, P8Struct, _i32,
                                   _i32, P8Struct, _vp]),
    "rtp_gn_apply_s2d": (C.c_int, [P8Struct, _i32, _i32, _vp, _vp, _vp, P8Struct, _vp]),
    "rtp_gn_bwd_reduce_s2d": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rtp_gn_bwd_apply_s2d": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, P8Struct, _i32,
                                       _i32, P8Struct, _vp]),
    "rtp_wgrad_pw_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "rtp_wgrad_pw_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_wgrad_pw": (C.c_int, [P8Struct, P8Struct, _i32, _vp, _vp, C.POINTER(C.c_int32), _vp]),
    "rtp_wgrad_pw_reduce": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rtp_reg_head_bwd_sparse_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_reg_head_bwd_sparse": (C.c_int, [P8Struct, P8Struct, _vp, _i32, _vp, _i32, _i32, P8Struct, _vp, _i32, _vp, _i32, _vp, _vp]),
    "rtp_reg_head_bwd_sparse_prezeroed": (C.c_int, [P8Struct, P8Struct, _vp, _i32, _vp, _i32, _i32, P8Struct, _vp, _i32, _vp, _i32, _vp,
                                                    _vp]),
    "rtp_zero_chunks": (C.c_int, [P8Struct, _vp]),
    "rtp_wgrad_pw_bias_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_wgrad_pw_bias": (C.c_int, [P8Struct, P8Struct, _i32, _vp, _vp, _vp, C.POINTER(C.c_int32), _vp]),
    "rtp_wgrad_pw_bias_reduce": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "rtp_wgrad_s2d_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "rtp_wgrad_s2d_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_wgrad_s2d": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, C.POINTER(C.c_int32), _vp]),
    "rtp_wgrad_s2d_reduce": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp]),
    "rtp_weight_s2d_expand": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "rtp_weight_s2d_fold": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "rtp_s2d_fold_weights": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "rtp_s2d_border_bias": (C.c_int, [_vp, P8Struct, _i32, _vp]),
    "rtp_s2d_box_sums_workspace_bytes": (C.c_int64, [_i32, _i32]),
    "rtp_s2d_fold_wgrad": (C.c_int, [P8Struct, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rtp_fuse_sum": (C.c_int, [C.POINTER(FuseDesc), _vp]),
    "rtp_conat_supported": (C.c_int, [C.POINTER(ConatDesc)]),
    "rtp_conat_fwd": (C.c_int, [C.POINTER(ConatDesc), _vp]),
    "rtp_upsample_bwd_workspace_bytes": (C.c_int64, [P8Struct, P8Struct, _i32]),
    "rtp_upsample_bwd": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp]),
    "rtp_grad_add": (C.c_int, [P8Struct, P8Struct, P8Struct, _i32, _i32, _vp]),
    "rtp_channel_sum": (C.c_int, [P8Struct, _i32, _vp, _i32, _vp, _vp]),
    "rtp_stem_fwd": (C.c_int, [P8Struct, _vp, _vp, _i32, P8Struct, _vp]),
    "rtp_stem_bwd": (C.c_int, [P8Struct, P8Struct, _i32, _vp, _vp, _i32, _vp, _vp]),
    "rtp_head_loss_workspace_bytes": (C.c_int64, [_i32, _i32, _i32, _i32, _i32]),
    "rtp_head_loss": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _f32, _vp,
                                P8Struct, P8Struct, _vp, _vp]),
    "rtp_head_loss_flags": (C.c_int, [P8Struct, P8Struct, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _f32, _vp,
                                      P8Struct, P8Struct, _i32, _vp, _vp]),
    "rtp_decode": (C.c_int, [P8Struct, P8Struct, _i32, _i32, C.POINTER(_f32), C.POINTER(_f32), _vp, _vp, _vp, _vp]),
    "rtp_dcn_fwd": (C.c_int, [_vp, _vp, _vp, _vp] + [_i32] * 11 + [_vp]),
    "rtp_dcn_bwd_input": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp] + [_i32] * 11 + [_vp]),
    "rtp_dcn_bwd_weight": (C.c_int, [_vp, _vp, _vp, _vp] + [_i32] * 11 + [_f32, _vp]),
    "rtp_pjpe": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "rtp_pjpe_seq_mean": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rtp_dcn_sample_p8": (C.c_int, [_vp, _vp, _vp, P8Struct] + [_i32] * 10 + [_vp]),
    "rtp_dcn_col2im_p8": (C.c_int, [_vp, _vp, _vp, P8Struct, _vp, _vp, _vp] + [_i32] * 10 + [_vp]),
    "rtp_dcn_bias_grad": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "rtp_mdcn_fwd": (C.c_int, [_vp] * 6 + [_i32] * 11 + [_vp]),
    "rtp_mdcn_bwd_input": (C.c_int, [_vp] * 8 + [_i32] * 11 + [_vp]),
    "rtp_mdcn_bwd_weight": (C.c_int, [_vp] * 6 + [_i32] * 11 + [_f32, _vp]),
    "rtp_scale_f32": (C.c_int, [_vp, _i64, _f32, _vp]),
    "rtp_adam_workspace_bytes": (C.c_int64, []),
    "rtp_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp, _vp, _vp]),
    "rtp_adam_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "rtp_assign_targets": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(C.c_double), C.POINTER(_f32), _vp,
                                     _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class RtpError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every prototype.  Raises if the .so has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtpError("librtpose_b200.so not built (%s missing): run `python -m rtpose_b200.build`; "
                       "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().rtp_last_error()
        raise RtpError("%s failed (rc=%d): %s" % (what or "rtp call", rc, msg.decode() if msg else "?"))


def require_device():
    import torch

    if not torch.cuda.is_available():
        raise RtpError("rtpose_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    if not load().rtp_device_ok():
        raise RtpError("rtpose_b200 kernels are built for sm_100a only; current device is not compute capability 10.x")
    setup_device()


_configured_devices = set()


def setup_device():
    """Once per device: one L1 / shared-memory preference for every kernel of the process (rtp_set_shared_carveout), so the
    small streaming kernels can become resident beside the persistent tensor-core CTAs.  RTP_NO_CARVEOUT=1 keeps the default."""
    import torch

    dev = torch.cuda.current_device()
    if dev in _configured_devices:
        return
    _configured_devices.add(dev)
    mode = int(os.environ.get("RTP_CARVEOUT", "0"))  # 0: defaults, 1: device-wide prefer-shared, 2: GroupNorm / finalize kernels only
    if mode:
        if load().rtp_set_shared_carveout(mode) != 0:
            raise RtpError(load().rtp_last_error().decode())


# kernels launched per C-ABI call (for the bench's `gpu_launches` claim)
LAUNCHES = {"rtp_pack_ncdhw": 1, "rtp_unpack_ncdhw": 1, "rtp_ingest_pack": 1, "rtp_weight_pack": 1,
            "rtp_weight_pack_k3s1": 1, "rtp_weight_pack_k3s1_window": 1, "rtp_weight_pack_batch": 1, "rtp_conv": 1, "rtp_conv_k3s1": 1, "rtp_wgrad": 1, "rtp_wgrad_reduce": 1,
            "rtp_gn_sums": 2, "rtp_gn_finalize": 1, "rtp_gn_apply": 1, "rtp_gn_bwd_reduce": 2, "rtp_gn_bwd_apply": 1,
            "rtp_fuse_sum": 1, "rtp_upsample_bwd": 2, "rtp_grad_add": 1, "rtp_channel_sum": 2, "rtp_stem_fwd": 1,
            "rtp_stem_bwd": 2, "rtp_head_loss": 2, "rtp_head_loss_flags": 2, "rtp_decode": 1, "rtp_scale_f32": 1, "rtp_dcn_fwd": 1,
            "rtp_dcn_bwd_input": 1, "rtp_dcn_bwd_weight": 1, "rtp_mdcn_fwd": 1, "rtp_mdcn_bwd_input": 1, "rtp_mdcn_bwd_weight": 2, "rtp_adam_step": 2, "rtp_adam_step_dev": 2,
            "rtp_assign_targets": 2, "rtp_wgrad_k3s1": 1, "rtp_wgrad_k3s1_reduce": 1, "rtp_conv_pw": 1, "rtp_gn_apply_s2d": 1, "rtp_gn_bwd_reduce_s2d": 2,
            "rtp_gn_bwd_apply_s2d": 1, "rtp_conv_k3s1_stat_finalize": 1, "rtp_conv_multi": 1, "rtp_gn_stats": 2,
            "rtp_wgrad_s2d": 1, "rtp_wgrad_s2d_reduce": 1, "rtp_s2d_fold_weights": 2, "rtp_s2d_border_bias": 1, "rtp_s2d_fold_wgrad": 3, "rtp_wgrad_pw": 1, "rtp_wgrad_pw_reduce": 1, "rtp_wgrad_pw_bias": 1, "rtp_wgrad_pw_bias_reduce": 1, "rtp_wgrad_pw_bias_workspace_bytes": 0,
            "rtp_conat_fwd": 1, "rtp_conat_supported": 0, "rtp_s2d_box_sums_workspace_bytes": 0,
            "rtp_npy_probe": 0, "rtp_npy_read_roi_slab": 0, "rtp_set_shared_carveout": 0, "rtp_reg_head_bwd_sparse": 3, "rtp_reg_head_bwd_sparse_prezeroed": 2, "rtp_zero_chunks": 1, "rtp_active_units": 1, "rtp_wgrad_k3s1_units": 1,
            "rtp_reg_head_bwd_sparse_workspace_bytes": 0}  # host-only file readers
launch_count = 0
call_counts = {}  # C-ABI entry point -> number of calls (tests assert which kernel path a shape really took)


def call(name, *args):
    global launch_count
    launch_count += LAUNCHES.get(name, 1)
    call_counts[name] = call_counts.get(name, 0) + 1
    check(getattr(load(), name)(*args), name)
