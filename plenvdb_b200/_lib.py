"""ctypes binding of libplenvdb_b200.so (the C-ABI declared in include/plenvdb_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every call
that returns a non-zero status raises ``PvdbError`` with the library's message.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libplenvdb_b200.so")


class PvdbError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libplenvdb_b200.so is missing (%s). Build it with `python -m plenvdb_b200.build` "
        "(nvcc, sm_100a); there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_f32p = C.c_void_p   # device or host pointers are passed as raw addresses
c_ptr = C.c_void_p


class pvdb_tree(C.Structure):
    _fields_ = [
        ("n_upper", C.c_int32), ("n_lower", C.c_int32), ("n_leaf", C.c_int32), ("reserved", C.c_int32),
        ("root_key0", C.c_uint64),
        ("root_keys", c_ptr), ("upper_child", c_ptr), ("lower_child", c_ptr), ("leaf_origin", c_ptr),
        ("leaf_mask", c_ptr),
    ]


class pvdb_train_cfg(C.Structure):
    _fields_ = [
        ("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3),
        ("reso", C.c_int32 * 3), ("mask_reso", C.c_int32 * 3),
        ("mask_scale", C.c_float * 3), ("mask_shift", C.c_float * 3),
        ("near", C.c_float), ("far", C.c_float), ("stepdist", C.c_float), ("act_shift", C.c_float),
        ("interval", C.c_float), ("fast_color_thres", C.c_float), ("bg", C.c_float),
        ("weight_main", C.c_float), ("weight_entropy_last", C.c_float), ("weight_rgbper", C.c_float),
        ("den_stepsz", C.c_float), ("k0_stepsz", C.c_float), ("eps", C.c_float), ("beta0", C.c_float),
        ("beta1", C.c_float),
        ("den_mode", C.c_int32), ("k0_mode", C.c_int32),
        ("net_lr", C.c_float), ("net_step", C.c_int32),
        ("k0_dim", C.c_int32), ("net_width", C.c_int32), ("use_tensor_cores", C.c_int32),
        ("n_rays_global", C.c_int32), ("parity_counts", C.c_int32),
    ]


class pvdb_train_bufs(C.Structure):
    _fields_ = [
        ("tree", C.POINTER(pvdb_tree)),
        ("den", c_ptr), ("den_grad", c_ptr), ("den_m", c_ptr), ("den_v", c_ptr),
        ("k0", c_ptr), ("k0_grad", c_ptr), ("k0_m", c_ptr), ("k0_v", c_ptr),
        ("occ_fine", c_ptr), ("occ_coarse", c_ptr),
        ("net", c_ptr), ("net_grad", c_ptr), ("net_m", c_ptr), ("net_v", c_ptr),
        ("t_min", c_ptr), ("t_max", c_ptr), ("n_steps", c_ptr),
        ("cnt_mask", c_ptr), ("cnt_alpha", c_ptr), ("cnt_keep", c_ptr), ("cnt_alpha_full", c_ptr),
        ("off_alpha", c_ptr), ("off_keep", c_ptr),
        ("alphainv_last", c_ptr), ("rgb_marched", c_ptr), ("grad_last", c_ptr),
        ("cap_alpha", C.c_int64), ("cap_keep", C.c_int64),
        ("s_ray", c_ptr), ("s_step", c_ptr),
        ("s_xyz", c_ptr), ("s_density", c_ptr), ("s_alpha", c_ptr), ("s_T", c_ptr), ("s_weight", c_ptr), ("s_gden", c_ptr),
        ("k_sample", c_ptr), ("k_ray", c_ptr),
        ("k_xyz", c_ptr), ("k_feat", c_ptr), ("k_rgb", c_ptr), ("k_gw", c_ptr),
        ("k_h0", c_ptr), ("k_h1", c_ptr), ("k_x", c_ptr), ("k_dh0", c_ptr), ("k_dh1", c_ptr), ("k_mask", c_ptr),
        ("k_corner", c_ptr), ("net_img", c_ptr), ("net_partial", c_ptr),
        ("march_scratch", c_ptr), ("scratch_rays", C.c_int32), ("scratch_per_ray", C.c_int32),
        ("den_touched", c_ptr), ("k0_touched", c_ptr), ("den_touched_list", c_ptr), ("k0_touched_list", c_ptr),
        ("counters", c_ptr), ("loss", c_ptr),
        ("den_perlr", c_ptr),
        ("ll_cnt", c_ptr), ("ll_off", c_ptr), ("ll_cur", c_ptr), ("ll_list", c_ptr), ("ll_items", c_ptr), ("k_dx", c_ptr),
        ("step_scalars", c_ptr), ("ray_pe", c_ptr),
    ]


class pvdb_render_cfg(C.Structure):
    _fields_ = [
        ("reso", C.c_int32 * 3), ("K", C.c_float * 9), ("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3),
        ("near", C.c_float), ("far", C.c_float), ("stepdist", C.c_float), ("act_shift", C.c_float),
        ("interval", C.c_float), ("fast_color_thres", C.c_float), ("bg", C.c_float),
        ("inverse_y", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("dcol", C.c_int32), ("dpe", C.c_int32), ("dhid", C.c_int32), ("dout", C.c_int32),
        ("use_tensor_cores", C.c_int32),
    ]


class pvdb_render_bufs(C.Structure):
    _fields_ = [
        ("idx_tree", C.POINTER(pvdb_tree)), ("idx_plane", c_ptr), ("dendata", c_ptr), ("coldata", c_ptr),
        ("w0", c_ptr), ("b0", c_ptr), ("w1", c_ptr), ("b1", c_ptr), ("w2", c_ptr), ("b2", c_ptr),
        ("n_samples", c_ptr), ("i_starts", c_ptr), ("tmins", c_ptr), ("tmaxs", c_ptr), ("scan_tmp", c_ptr),
        ("cap_samples", C.c_int64),
        ("s_ray", c_ptr), ("s_weight", c_ptr), ("s_feat", c_ptr), ("s_rgb", c_ptr), ("counters", c_ptr),
        ("w_img", c_ptr), ("active_list", c_ptr), ("skip_bits", c_ptr),
        ("px_scratch", c_ptr), ("fallback_list", c_ptr), ("px_entries", C.c_int32), ("reserved", C.c_int32),
    ]


class pvdb_dp_peers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_leaf", C.c_int32), ("cap_leaves", C.c_int32),
                ("base", C.c_void_p * 8)]


class pvdb_frame_peers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("root", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("reserved", C.c_int32), ("base", C.c_void_p * 8)]


_TP = C.POINTER(pvdb_tree)
_i, _i64, _f = C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); None restype = int status checked by `call`
_SIGS = {
    "pvdb_last_error": (C.c_char_p, []),
    "pvdb_abi_version": (C.c_int, []),
    "pvdb_last_launch_count": (C.c_int, []),
    "pvdb_topo_create_dense": (c_ptr, [_i, _i, _i]),
    "pvdb_topo_create_from_mask": (c_ptr, [c_ptr, _i, _i, _i]),
    "pvdb_topo_destroy": (None.__class__, [c_ptr]),
    "pvdb_topo_counts": (None, [c_ptr, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pvdb_topo_export": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_sample_forward": (None, [_TP, c_ptr, _i, c_ptr, c_ptr, c_ptr, _i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_sample_backward": (None, [_TP, c_ptr, _i, c_ptr, c_ptr, c_ptr, c_ptr, _i64, c_ptr]),
    "pvdb_sample_forward_host": (None, [_TP, c_ptr, _i, c_ptr, c_ptr, c_ptr, _i64, c_ptr, c_ptr]),
    "pvdb_sample_backward_host": (None, [_TP, c_ptr, _i, c_ptr, c_ptr, c_ptr, c_ptr, _i64, c_ptr]),
    "pvdb_sample_nearest": (None, [_TP, c_ptr, _i, c_ptr, c_ptr, c_ptr, _i64, c_ptr, c_ptr]),
    "pvdb_adam_stepsize": (C.c_float, [_f, _f, _f, _i]),
    "pvdb_adam_step": (None, [_TP, c_ptr, c_ptr, c_ptr, c_ptr, _i, _i, _f, _f, _f, _f, c_ptr, c_ptr]),
    "pvdb_zero_grad": (None, [_TP, c_ptr, _i, c_ptr]),
    "pvdb_copy_from_dense": (None, [_TP, c_ptr, _i, c_ptr, _i, _i, _i, c_ptr]),
    "pvdb_copy_to_dense": (None, [_TP, c_ptr, _i, c_ptr, _i, _i, _i, c_ptr]),
    "pvdb_set_values_on_by_mask": (None, [_TP, c_ptr, c_ptr, _f, _i, _i, _i, c_ptr]),
    "pvdb_infer_t_minmax": (None, [c_ptr, c_ptr, c_ptr, c_ptr, _f, _f, _i, c_ptr, c_ptr, c_ptr]),
    "pvdb_infer_n_samples": (None, [c_ptr, c_ptr, c_ptr, _f, _i, c_ptr, c_ptr]),
    "pvdb_infer_ray_start_dir": (None, [c_ptr, c_ptr, c_ptr, _i, c_ptr, c_ptr, c_ptr]),
    "pvdb_sample_pts_count": (None, [c_ptr, c_ptr, c_ptr, c_ptr, _f, _f, _f, _i, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                     c_ptr, c_ptr]),
    "pvdb_sample_pts_fill": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _f, _i, _i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_maskcache_lookup": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _i, _i, _i, _i64, c_ptr]),
    "pvdb_raw2alpha": (None, [c_ptr, _f, _f, _i64, c_ptr, c_ptr, c_ptr]),
    "pvdb_raw2alpha_backward": (None, [c_ptr, c_ptr, _f, _i64, c_ptr, c_ptr]),
    "pvdb_alpha2weight": (None, [c_ptr, c_ptr, _i64, _i, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_alpha2weight_backward": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _i, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_dense_adam": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _i64, _i, _i, _f, _f, _f, _f, c_ptr]),
    "pvdb_occ_build": (None, [c_ptr, _i, _i, _i, c_ptr, c_ptr, c_ptr]),
    "pvdb_dp_pack": (None, [C.POINTER(pvdb_train_bufs), c_ptr, c_ptr, c_ptr, c_ptr, _i64, c_ptr]),
    "pvdb_dp_unpack": (None, [C.POINTER(pvdb_train_bufs), c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_dp_symm_bytes": (C.c_size_t, [_i, _i]),
    "pvdb_dp_symm_alloc": (None, [C.c_size_t, C.POINTER(C.c_void_p), c_ptr]),
    "pvdb_dp_symm_open": (None, [c_ptr, C.POINTER(C.c_void_p)]),
    "pvdb_dp_symm_close": (None, [c_ptr]),
    "pvdb_dp_symm_free": (None, [c_ptr]),
    "pvdb_dp_grad_planes": (None, [C.POINTER(pvdb_dp_peers), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "pvdb_dp_symm_error": (None, [C.POINTER(pvdb_dp_peers), C.POINTER(C.c_int32)]),
    "pvdb_dp_exchange": (None, [C.POINTER(pvdb_dp_peers), C.POINTER(pvdb_train_bufs), C.c_uint32, c_ptr]),
    "pvdb_dp_exchange_tiles": (None, [C.POINTER(pvdb_dp_peers), C.POINTER(pvdb_train_bufs), C.c_uint32, c_ptr]),
    "pvdb_dp_exchange_net": (None, [C.POINTER(pvdb_dp_peers), C.POINTER(pvdb_train_bufs), C.c_uint32, c_ptr]),
    "pvdb_train_step_dp": (None, [C.POINTER(pvdb_train_cfg), C.POINTER(pvdb_train_bufs), C.POINTER(pvdb_dp_peers), C.c_uint32,
                                  c_ptr, c_ptr, c_ptr, c_ptr, _i, c_ptr]),
    "pvdb_resample_trilinear": (None, [_TP, c_ptr, _i, _i, _i, _i, _TP, c_ptr, _i, _i, _i, c_ptr]),
    "pvdb_plane_remap": (None, [_TP, c_ptr, _TP, c_ptr, _i, c_ptr]),
    "pvdb_occupancy_update": (None, [_TP, c_ptr, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _f, _f, _f, c_ptr, _i, _i, _i, c_ptr, c_ptr]),
    "pvdb_total_variation_add_grad": (None, [_TP, c_ptr, c_ptr, _i, _i, _i, _i, _f, _f, _f, _i, c_ptr]),
    "pvdb_debug_set_run_skip": (None, [_i]),
    "pvdb_debug_set_leaf_local": (None, [_i]),
    "pvdb_debug_stamps_fetch": (C.c_int, [c_ptr]),
    "pvdb_debug_set_render_lanes": (None, [_i]),
    "pvdb_stage_rays": (None, [c_ptr, c_ptr, c_ptr, c_ptr, _i, c_ptr, c_ptr]),
    "pvdb_dense_adam_stepsize_host": (C.c_float, [_f, _f, _f, _i]),
    "pvdb_profile_enable": (None, [_i]),
    "pvdb_profile_fetch": (C.c_int, [_i, c_ptr, c_ptr]),
    "pvdb_rays_hit_mask": (None, [C.POINTER(pvdb_train_cfg), C.POINTER(pvdb_train_bufs), c_ptr, c_ptr, _i, c_ptr, c_ptr]),
    "pvdb_render_rows": (None, [C.POINTER(pvdb_render_cfg), C.POINTER(pvdb_render_bufs), c_ptr, _i, _i, c_ptr, c_ptr]),
    "pvdb_interleaved_rows": (C.c_int, [_i, _i, _i, _i]),
    "pvdb_render_rows_interleaved": (None, [C.POINTER(pvdb_render_cfg), C.POINTER(pvdb_render_bufs), c_ptr, _i, _i, _i, c_ptr, c_ptr,
                                            c_ptr]),
    "pvdb_frame_symm_bytes": (C.c_size_t, [_i, _i]),
    "pvdb_render_frame_sharded": (None, [C.POINTER(pvdb_render_cfg), C.POINTER(pvdb_render_bufs), C.POINTER(pvdb_frame_peers), c_ptr,
                                         _i, C.c_uint32, c_ptr, c_ptr]),
    "pvdb_frame_ptr": (None, [C.POINTER(pvdb_frame_peers), C.c_uint32, C.POINTER(C.c_void_p)]),
    "pvdb_frame_copy": (None, [C.POINTER(pvdb_frame_peers), C.c_uint32, c_ptr, c_ptr]),
    "pvdb_frame_error": (None, [C.POINTER(pvdb_frame_peers), C.POINTER(C.c_int32)]),
    "pvdb_render_block_bits_words": (C.c_size_t, [_i, _i, _i]),
    "pvdb_render_block_bits": (None, [_TP, _i, _i, _i, c_ptr, c_ptr]),
    "pvdb_merge_gather": (None, [_TP, c_ptr, c_ptr, _i, c_ptr, _i, _i, _i, c_ptr, c_ptr, c_ptr]),
    "pvdb_train_step": (None, [C.POINTER(pvdb_train_cfg), C.POINTER(pvdb_train_bufs), c_ptr, c_ptr, c_ptr, c_ptr, _i, _i,
                               c_ptr]),
    # opaque handles (csrc/handles.cu)
    "pvdb_grid_create": (c_ptr, [_i, _i, _i, _i, c_ptr]),
    "pvdb_grid_destroy": (type(None), [c_ptr]),
    "pvdb_grid_info": (None, [c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_grid_tree": (c_ptr, [c_ptr]),
    "pvdb_grid_values": (c_ptr, [c_ptr]),
    "pvdb_grid_grad": (c_ptr, [c_ptr]),
    "pvdb_grid_copy_from_dense": (None, [c_ptr, c_ptr]),
    "pvdb_grid_copy_to_dense": (None, [c_ptr, c_ptr]),
    "pvdb_grid_forward": (None, [c_ptr, c_ptr, c_ptr, c_ptr, _i64, c_ptr]),
    "pvdb_grid_backward": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _i64]),
    "pvdb_grid_set_values_on_by_mask": (None, [c_ptr, c_ptr, _f]),
    "pvdb_opt_create": (c_ptr, [c_ptr, _f, _f, _f, _f]),
    "pvdb_opt_destroy": (type(None), [c_ptr]),
    "pvdb_opt_zero_grad": (None, [c_ptr]),
    "pvdb_opt_step": (None, [c_ptr, _i]),
    "pvdb_opt_update_lr": (None, [c_ptr, _f]),
    "pvdb_opt_set_pervoxel_lr": (None, [c_ptr, c_ptr]),
    "pvdb_opt_get": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_opt_set": (None, [c_ptr, C.c_int32, _f, _f, _f, _f]),
    "pvdb_opt_exp_avg": (c_ptr, [c_ptr]),
    "pvdb_opt_exp_avg_sq": (c_ptr, [c_ptr]),
    "pvdb_renderer_create": (c_ptr, [_i, _i, _i, _i]),
    "pvdb_renderer_destroy": (type(None), [c_ptr]),
    "pvdb_renderer_load_data": (None, [c_ptr, c_ptr, c_ptr, _i64, c_ptr, _i, _i, _i]),
    "pvdb_renderer_load_params": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_renderer_set_scene": (None, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pvdb_renderer_set_kwargs": (None, [c_ptr, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i]),
    "pvdb_renderer_input_c2w": (None, [c_ptr, c_ptr]),
    "pvdb_renderer_render": (None, [c_ptr, c_ptr, c_ptr]),
    "pvdb_renderer_frame": (c_ptr, [c_ptr]),
    "pvdb_renderer_counters": (None, [c_ptr, c_ptr]),
}

DECLARED_SYMBOLS = sorted(_SIGS)
_missing = []
for _name, (_res, _args) in _SIGS.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        _missing.append(_name)
        continue
    _fn.argtypes = _args
    if _res is None:
        _fn.restype = C.c_int
    elif _res is type(None):
        _fn.restype = None
    else:
        _fn.restype = _res
MISSING_SYMBOLS = _missing


def last_error():
    return lib.pvdb_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Invoke a status-returning entry point; raise PvdbError on failure."""
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PvdbError("%s failed (status %d): %s" % (name, rc, last_error()))


def ptr(t):
    """Raw address of a torch tensor (device or host) or None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class _DevicePointer:
    """Exposes a raw device allocation (not owned by torch) through __cuda_array_interface__."""

    def __init__(self, address, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(address), False), "version": 2,
                                         "strides": None}


def tensor_from_ptr(address, shape, device):
    """float32 CUDA tensor VIEW of `shape` over a device allocation made by the library (e.g. a symmetric block); the caller
    keeps the allocation alive."""
    import torch
    return torch.as_tensor(_DevicePointer(address, shape), device=device)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


PROFILING = False      # per-kernel event profiling is on: the fused calls run serially on one stream (and never as a graph replay)


def profile_enable(on=True):
    global PROFILING
    PROFILING = bool(on)
    call("pvdb_profile_enable", int(bool(on)))


def profile_fetch(max_segments=48):
    """[(kernel name, ms)] of the last fused call issued with profiling enabled."""
    ms = (C.c_float * max_segments)()
    names = C.create_string_buffer(max_segments * 32)
    n = lib.pvdb_profile_fetch(max_segments, ms, names)
    return [(names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode(), float(ms[i])) for i in range(n)]
