"""Mirror of the reference's torch extensions ``render_utils_cuda`` and ``adam_upd_cuda``
(plenvdb/lib/cuda/render_utils.cpp:170-184, adam_upd.cpp): same function names, same tensor
contracts (contiguous CUDA tensors in, fresh CUDA tensors out, int64 ids), implemented by the C-ABI
ops of libplenvdb_b200 (csrc/render_utils.cu).  Callers: plenvdb/lib/dvgo.py:265, 288, 417, 432, 453,
463 and plenvdb/lib/grid.py:241.
"""
import torch

from . import _lib
from ._lib import call, current_stream, ptr


_FLOAT_OK = (torch.float32,)
_OTHER_OK = (torch.int64, torch.bool, torch.uint8)


def _check(*ts):
    """CHECK_INPUT (render_utils.cpp:46-48) plus the dtype contract: the reference dispatches on float / double
    (AT_DISPATCH_FLOATING_TYPES); these kernels are float32 only (ids int64, masks bool), so anything else is refused
    instead of being reinterpreted."""
    for t in ts:
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("expected a contiguous CUDA tensor")
        if t.dtype not in _FLOAT_OK and t.dtype not in _OTHER_OK:
            raise RuntimeError("render_utils_cuda (B200): float32 tensors only (got %s); the reference's float64 dispatch is not built" % t.dtype)


def _out_of_scope(name, users):
    def fn(*a, **k):
        raise NotImplementedError("render_utils_cuda.%s serves %s only, which is outside the ported hot path (SURVEY.md 8b, B2)" % (name, users))
    fn.__name__ = name
    return fn


# the four ops of the extension that only the unbounded / MPI / non-uniform model variants call (dmpigo.py:240, dbvgo.py:235,243,
# dvgo.py's raw2alpha_nonuni): present so that attribute lookups succeed, loud when called
sample_ndc_pts_on_rays = _out_of_scope("sample_ndc_pts_on_rays", "DirectMPIGO (dmpigo.py)")
sample_bg_pts_on_rays = _out_of_scope("sample_bg_pts_on_rays", "DirectBiVoxGO (dbvgo.py)")
raw2alpha_nonuni = _out_of_scope("raw2alpha_nonuni", "the non-uniform step models (dbvgo.py / dcvgo.py)")
raw2alpha_nonuni_backward = _out_of_scope("raw2alpha_nonuni_backward", "the non-uniform step models (dbvgo.py / dcvgo.py)")


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    _check(rays_o, rays_d, xyz_min, xyz_max)
    n = rays_o.shape[0]
    t_min = torch.empty(n, dtype=rays_o.dtype, device=rays_o.device)
    t_max = torch.empty_like(t_min)
    call("pvdb_infer_t_minmax", ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max), float(near), float(far), n, ptr(t_min),
         ptr(t_max), current_stream())
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    _check(rays_d, t_min, t_max)
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=t_min.device)
    call("pvdb_infer_n_samples", ptr(rays_d), ptr(t_min), ptr(t_max), float(stepdist), n, ptr(out), current_stream())
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    _check(rays_o, rays_d, t_min)
    n = rays_o.shape[0]
    start, direc = torch.empty_like(rays_o), torch.empty_like(rays_o)
    call("pvdb_infer_ray_start_dir", ptr(rays_o), ptr(rays_d), ptr(t_min), n, ptr(start), ptr(direc), current_stream())
    return [start, direc]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """-> [rays_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max] (render_utils_kernel.cu:196-242)."""
    _check(rays_o, rays_d, xyz_min, xyz_max)
    n = rays_o.shape[0]
    dev = rays_o.device
    t_min = torch.empty(n, dtype=torch.float32, device=dev)
    t_max = torch.empty_like(t_min)
    n_steps = torch.empty(n, dtype=torch.int64, device=dev)
    cumsum = torch.empty_like(n_steps)
    start, direc = torch.empty_like(rays_o), torch.empty_like(rays_o)
    call("pvdb_sample_pts_count", ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max), float(near), float(far),
         float(stepdist), n, ptr(t_min), ptr(t_max), ptr(n_steps), ptr(cumsum), ptr(start), ptr(direc), current_stream())
    total = int(cumsum[-1].item()) if n else 0          # the reference's .item() sync (:212)
    rays_pts = torch.empty((total, 3), dtype=torch.float32, device=dev)
    mask_outbbox = torch.empty(total, dtype=torch.bool, device=dev)
    ray_id = torch.empty(total, dtype=torch.int64, device=dev)
    step_id = torch.empty(total, dtype=torch.int64, device=dev)
    call("pvdb_sample_pts_fill", ptr(start), ptr(direc), ptr(xyz_min), ptr(xyz_max), ptr(cumsum), float(stepdist), n, total,
         ptr(rays_pts), ptr(mask_outbbox), ptr(ray_id), ptr(step_id), current_stream())
    return [rays_pts, mask_outbbox, ray_id, step_id, n_steps, t_min, t_max]


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    _check(world, xyz, xyz2ijk_scale, xyz2ijk_shift)
    assert world.dtype == torch.bool and world.dim() == 3
    n = xyz.shape[0]
    out = torch.zeros(n, dtype=torch.bool, device=xyz.device)
    call("pvdb_maskcache_lookup", ptr(world), ptr(xyz), ptr(out), ptr(xyz2ijk_scale), ptr(xyz2ijk_shift), world.shape[0],
         world.shape[1], world.shape[2], n, current_stream())
    return out


def raw2alpha(density, shift, interval):
    _check(density)
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    call("pvdb_raw2alpha", ptr(density), float(shift), float(interval), density.numel(), ptr(exp_d), ptr(alpha),
         current_stream())
    return [exp_d, alpha]


def raw2alpha_backward(exp_d, grad_back, interval):
    _check(exp_d, grad_back)
    grad = torch.empty_like(exp_d)
    call("pvdb_raw2alpha_backward", ptr(exp_d), ptr(grad_back), float(interval), exp_d.numel(), ptr(grad), current_stream())
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    _check(alpha, ray_id)
    if ray_id.dtype != torch.int64 or alpha.dtype != torch.float32:
        raise RuntimeError("alpha2weight: alpha float32 and ray_id int64 expected")
    n_pts = alpha.shape[0]
    dev = alpha.device
    weight = torch.zeros_like(alpha)
    T = torch.ones_like(alpha)
    alphainv_last = torch.ones(n_rays, dtype=alpha.dtype, device=dev)
    i_start = torch.zeros(n_rays, dtype=torch.int64, device=dev)
    i_end = torch.zeros(n_rays, dtype=torch.int64, device=dev)
    call("pvdb_alpha2weight", ptr(alpha), ptr(ray_id), n_pts, int(n_rays), ptr(weight), ptr(T), ptr(alphainv_last),
         ptr(i_start), ptr(i_end), current_stream())
    return [weight, T, alphainv_last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    _check(alpha, weight, T, alphainv_last, i_start, i_end, grad_weights, grad_last)
    grad = torch.zeros_like(alpha)
    call("pvdb_alpha2weight_backward", ptr(alpha), ptr(weight), ptr(T), ptr(alphainv_last), ptr(i_start), ptr(i_end),
         int(n_rays), ptr(grad_weights), ptr(grad_last), ptr(grad), current_stream())
    return grad


# ---- adam_upd_cuda (plenvdb/lib/cuda/adam_upd.cpp; caller plenvdb/lib/masked_adam.py:149-159)
def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _check(param, grad, exp_avg, exp_avg_sq)
    call("pvdb_dense_adam", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), None, param.numel(), 0, int(step),
         float(beta1), float(beta2), float(lr), float(eps), current_stream())


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _check(param, grad, exp_avg, exp_avg_sq)
    call("pvdb_dense_adam", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), None, param.numel(), 1, int(step),
         float(beta1), float(beta2), float(lr), float(eps), current_stream())


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _check(param, grad, exp_avg, exp_avg_sq, perlr)
    call("pvdb_dense_adam", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(perlr), param.numel(), 2, int(step),
         float(beta1), float(beta2), float(lr), float(eps), current_stream())
