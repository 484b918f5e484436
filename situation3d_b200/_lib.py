"""ctypes binding of libpn2_b200.so (C ABI: include/pn2_b200.h).

PyTorch is plumbing here: tensors supply device pointers and the current stream.
The library must exist -- there is no fallback implementation.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PN2_B200_LIB") or os.path.join(_HERE, "libpn2_b200.so")   # PN2_B200_LIB: an A/B build (build.py)

PN2_OK = 0


class Pn2Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libpn2_b200.so is not built (%s). Run `python -m situation3d_b200.build`; "
            "this package has no CPU or PyTorch fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_p = c_void_p
_i = c_int
_f = c_float
_ll = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/pn2_b200.h declaration by declaration
_PROTOTYPES = {
    "pn2_version": (_i, []),
    "pn2_error_string": (c_char_p, [_i]),
    "pn2_last_cuda_error": (c_char_p, []),
    "pn2_device_check": (_i, []),
    "pn2_launch_count": (ctypes.c_ulonglong, []),
    "pn2_sm_partition_create": (_i, [_i, POINTER(_p)]),
    "pn2_sm_partition_sms": (_i, [_p, _i]),
    "pn2_sm_partition_stream_create": (_i, [_p, _i, POINTER(_p)]),
    "pn2_stream_sm_count": (_i, [_p]),
    "pn2_gather_points": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "pn2_gather_points_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "pn2_furthest_point_sampling_workspace_bytes": (c_size_t, [_i, _i, _i]),
    "pn2_furthest_point_sampling": (_i, [_i, _i, _i, _p, _p, c_size_t, _p, _p]),
    "pn2_furthest_point_sampling_xyz": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "pn2_furthest_point_sampling_rows": (_i, [_i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "pn2_furthest_point_sampling_xyz_ws": (_i, [_i, _i, _i, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_furthest_point_sampling_rows_ws": (_i, [_i, _i, _i, _p, _i, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_debug_fps_bucket_profile": (_i, [_i, _i, _i, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_debug_fps_profile": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "pn2_three_nn": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_three_interpolate": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_three_interpolate_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_ball_query": (_i, [_i, _i, _i, _f, _i, _p, _p, _p, _p]),
    "pn2_ball_query_workspace_bytes": (c_size_t, [_i, _i, _i, _i]),
    "pn2_ball_query_ws": (_i, [_i, _i, _i, _f, _i, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_furthest_point_sampling_workspace_bytes_mode": (c_size_t, [_i, _i, _i, _i]),
    "pn2_ball_query_grid_build": (_i, [_i, _i, _i, _f, _i, _p, _i, _p, c_size_t, _p]),
    "pn2_ball_query_grid_query": (_i, [_i, _i, _i, _f, _i, _p, _p, _p, c_size_t, _p]),
    "pn2_group_rows": (_i, [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _i, _p, _p]),
    "pn2_rows_bn_supported": (_i, [_ll, _i, _i]),
    "pn2_rows_bn_partials_bytes": (c_size_t, [_ll, _i]),
    "pn2_rows_bn_stats": (_i, [_ll, _i, _p, _p, POINTER(_i), _p]),
    "pn2_rows_bn_finalize": (_i, [_i, _i, _p, _ll, ctypes.c_double, _p, _p, _f, _p, _p, _p, _p, _p, _p]),
    "pn2_rows_bn_bwd_finalize": (_i, [_i, _i, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_rows_bn_relu_apply": (_i, [_ll, _i, _p, _p, _p, _p, _p]),
    "pn2_rows_bn_relu_pool": (_i, [_ll, _i, _i, _p, _p, _p, _p, _p, _p]),
    "pn2_rows_bn_relu_bwd_reduce": (_i, [_ll, _i, _p, _p, _p, _p, _p, POINTER(_i), _p]),
    "pn2_rows_bn_relu_bwd_apply": (_i, [_ll, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_rows_bn_relu_pool_bwd_reduce": (_i, [_ll, _i, _i, _p, _p, _p, _p, _p, _p, POINTER(_i), _p]),
    "pn2_rows_bn_relu_pool_bwd_apply": (_i, [_ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_group_points": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "pn2_group_points_grad": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "pn2_mlp_f32_image_bytes": (c_size_t, [_i, POINTER(_i)]),
    "pn2_mlp_f32_supported": (_i, [_i, POINTER(_i)]),
    "pn2_mlp_f32_pack": (_i, [_i, POINTER(_i), POINTER(_p), POINTER(_p), _p, _p]),
    "pn2_rows_from_channels": (_i, [_i, _i, _i, _p, _p, _p]),
    "pn2_sa_forward_f32": (_i, [_i, _i, _i, _i, _i, _p, _i, _i, _f, _p, _p, _p, _i, POINTER(_i), _p, _p, _p, _p]),
    "pn2_fp_forward_f32": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, POINTER(_i), _p, _p, _p, _p]),
    "pn2_sa_tc_row_elems": (_i, [_i]),
    "pn2_sa_tc_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "pn2_sa_tc_weight_image_bytes": (c_size_t, [_i, _i, _i, _i]),
    "pn2_sa_tc_pack_weights": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_sa_tc_pack_rows": (_i, [_i, _i, _i, _p, ctypes.c_longlong, _i, _p, _p]),
    "pn2_sa_tc_pack_channels": (_i, [_i, _i, _i, _p, _p, _p]),
    "pn2_sa_tc_forward": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_lin_tc_supported": (_i, [_i, _i, _i]),
    "pn2_lin_tc_weight_image_bytes": (c_size_t, [_i, _i]),
    "pn2_lin_tc_pack_weights": (_i, [_i, _i, _i, _i, _p, _i, _i, _p, _p]),
    "pn2_lin_tc_forward": (_i, [ctypes.c_longlong, _i, _i, _p, _i, _i, _p, _p, _p]),
    "pn2_fp_tc_supported": (_i, [_i, _i, _i, _i]),
    "pn2_fp_tc_weight_image_bytes": (c_size_t, [_i, _i, _i, _i]),
    "pn2_fp_tc_pack_weights": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "pn2_fp_tc_forward": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_fp_tc2_supported": (_i, [_i, _i, _i, _i, _i]),
    "pn2_fp_tc2_forward": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_debug_fp_tc2_profile": (_i, [_p]),
    "pn2_debug_sa_tc_profile": (_i, [_p]),
    "pn2_selftest_umma": (_i, [_i, _i, _p, _p, _p, _p]),
    "pn2_linear_gelu_tc_supported": (_i, [_i, _i]),
    "pn2_linear_gelu_tc_weight_image_bytes": (c_size_t, [_i, _i]),
    "pn2_linear_gelu_tc_pack_weights": (_i, [_i, _i, _p, _p, _p]),
    "pn2_linear_gelu_tc_forward": (_i, [ctypes.c_longlong, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_column_pool_max_voxels": (_i, []),
    "pn2_column_pool": (_i, [_i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "pn2_token_gather": (_i, [_i, _i, _i, _p, _p, _p, _p, _f, _f, _f, _p, _p, _p]),
    "pn2_voxel_pe": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _i, _p, _p]),
    "pn2_frustum_planes": (_i, [_i, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "pn2_points_in_frustum": (_i, [_i, _p, _p, _p, _p, _p, _p]),
    "pn2_compute_projection_workspace_bytes": (c_size_t, [_i, _i]),
    "pn2_compute_projection": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_project_workspace_bytes": (c_size_t, [_i, _i]),
    "pn2_project": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, c_size_t, _p]),
    "pn2_project_maxpool": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, c_size_t, _p]),
    "pn2_quaternions_to_rotation_matrices": (_i, [_i, _p, _p, _p]),
    "pn2_rotation_vectors_to_matrices": (_i, [_i, _p, _p, _p]),
    "pn2_situation_matrices": (_i, [_i, _p, _p, _p]),
    "pn2_reencode_forward": (_i, [_i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}

for _name, (_res, _args) in _PROTOTYPES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc, what):
    """Raise like the reference's AT_ASSERT / CUDA_CHECK_ERRORS would, but never exit()."""
    if rc != PN2_OK:
        msg = lib.pn2_error_string(rc).decode()
        if rc == -2:
            msg += ": " + lib.pn2_last_cuda_error().decode()
        raise Pn2Error("%s failed: %s" % (what, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def ptr_array(tensors):
    return (c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
