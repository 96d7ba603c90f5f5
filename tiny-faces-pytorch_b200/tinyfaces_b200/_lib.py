"""ctypes binding of libtinyfaces_b200.so (the C-ABI boundary, include/tinyfaces_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtinyfaces_b200.so")
_lib = None

c_i32, c_i64, c_f32, c_f64 = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double
c_u32, c_u64, c_vp, c_sz = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t

_SIGS = {
    "tf_version": (c_i32, []),
    "tf_debug_set": (c_i32, [c_i32, c_i32]),
    "tf_gemm_error_flag": (c_i32, [ctypes.POINTER(c_i32)]),
    "tf_nms_set_algorithm": (c_i32, [c_i32]),
    "tf_nms_workspace_bytes": (c_i32, [c_i64, c_i32, ctypes.POINTER(c_sz)]),
    "tf_nms": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "tf_nms_algo": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f64, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "tf_nms_sweep_stats": (c_i32, [c_i64, c_i32, c_vp, c_sz, ctypes.POINTER(c_i64), c_vp]),
    "tf_decode_workspace_bytes": (c_i32, [c_i64, c_i64, c_i64, ctypes.POINTER(c_sz)]),
    "tf_decode": (c_i32, [c_vp, c_vp, c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_i32, c_i32, c_i32, c_i32,
                          ctypes.POINTER(c_f64), c_f32, c_u32, c_u32, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                          c_f64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "tf_detloss_ohem": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i64, c_f32, c_vp]),
    "tf_detloss_fwd_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "tf_detloss_sample_workspace_bytes": (c_i32, [c_i32, ctypes.POINTER(c_sz)]),
    "tf_detloss_sample_device": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_i32, c_u64, c_vp, c_sz, c_vp]),
    "tf_detloss_sample_device_ctr": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_i32, c_u64, c_vp, c_vp, c_sz, c_vp]),
    "tf_model_create": (c_i32, [c_i32, ctypes.POINTER(c_vp)]),
    "tf_model_destroy": (c_i32, [c_vp]),
    "tf_model_num_params": (c_i32, [c_vp]),
    "tf_model_param_name": (ctypes.c_char_p, [c_vp, c_i32]),
    "tf_model_output_shape": (c_i32, [c_vp, c_i32, c_i32, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "tf_model_workspace_bytes": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(c_sz)]),
    "tf_model_forward": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(c_vp), c_i32, c_i32, c_f32, c_vp, c_vp,
                                 c_sz, c_vp]),
    "tf_model_backward": (c_i32, [c_vp, c_vp, ctypes.POINTER(c_vp), c_vp]),
    "tf_model_backward_ex": (c_i32, [c_vp, c_vp, ctypes.POINTER(c_vp), c_i32, ctypes.POINTER(c_vp), ctypes.POINTER(c_i32), c_vp]),
    "tf_model_get_tensor": (c_i32, [c_vp, ctypes.c_char_p, c_vp, c_i64, ctypes.POINTER(c_i32), c_vp]),
    "tf_model_upsample_offdiag": (c_i32, [c_vp, ctypes.POINTER(c_f32), c_vp]),
    "tf_targets_workspace_bytes": (c_i32, [c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(c_sz)]),
    "tf_heatmap_targets": (c_i32, [c_vp, c_i32, ctypes.POINTER(c_f64), c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f64, c_f64,
                                   c_vp, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "tf_sgd_step": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, ctypes.POINTER(c_i64), ctypes.POINTER(c_f32), ctypes.POINTER(c_f32),
                            c_f32, c_f32, c_vp, c_vp]),
    "tf_steplr_update": (c_i32, [c_vp, c_vp, c_i32, c_f32, c_i32, c_vp]),
    "tf_pyramid_workspace_bytes": (c_i32, [c_i32, c_i32, c_i32, ctypes.POINTER(c_sz)]),
    "tf_pyramid_level": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32,
                                 ctypes.POINTER(c_f32), ctypes.POINTER(c_f32), c_vp, c_vp, c_sz, c_vp]),
    "tf_bn_workspace_bytes": (c_i32, [ctypes.POINTER(c_sz)]),
    "tf_bn_train_fwd": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_f32, c_f32, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_sz, c_vp]),
    "tf_bn_eval_fwd": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_i32, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "tf_bn_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_sz, c_vp]),
    "tf_maxpool_fwd": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "tf_maxpool_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "tf_head_workspace_bytes": (c_i32, [c_i32, ctypes.POINTER(c_sz)]),
    "tf_head_upsample_add_fwd": (c_i32, [c_vp, c_vp, c_vp] + [c_i32] * 7 + [c_vp, c_vp, c_sz, c_vp]),
    "tf_head_upsample_add_bwd": (c_i32, [c_vp, c_vp] + [c_i32] * 7 + [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "tf_conv2d_nhwc": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "tf_conv2d_nhwc_res": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "tf_conv2d_nhwc_strided": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "tf_conv2d_wgrad_nhwc_strided": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "tf_conv_plan": (c_i32, [c_i32] * 11 + [ctypes.POINTER(c_i32)]),
    "tf_dgrad_s2_taps": (c_i32, [c_i32, c_i32, c_i32, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "tf_conv2d_dgrad_s2_nhwc": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "tf_conv2d_wgrad_nhwc": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError("libtinyfaces_b200.so is missing (%s): run `python tiny-faces-pytorch_b200/build.py`; "
                               "there is no CPU fallback" % _SO)
        l = ctypes.CDLL(_SO)
        l.tf_last_error_string.restype = ctypes.c_char_p
        l.tf_last_error_string.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name, None)
            if fn is None:
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS) + ["tf_last_error_string"]


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("tinyfaces_b200 %s failed (%d): %s" % (what, rc, lib().tf_last_error_string().decode()))


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s must be a CUDA tensor: the tinyfaces_b200 kernels have no CPU fallback" % name)
