"""Deterministic synthetic inputs of SURVEY.md §8(d): NeRF-Synthetic-shaped scenes, cameras, ray batches and
rgbnet weights.  Pure numpy (float32) so the same bits feed the CUDA path, the oracle and the reference
harnesses.  Shapes and scalars follow the reference's fine stage (plenvdb/configs/default.py:41-115,
plenvdb/lib/load_blender.py:29-34, plenvdb/lib/load_data.py:143-151, plenvdb/lib/dvgo.py:471-499).
"""
import math

import numpy as np

F32 = np.float32
NET_N = 22019   # 39*128+128 + 128*128+128 + 128*3+3


def _dilate(occ):
    """3^3 max-pool, padding 1 (F.max_pool3d(kernel_size=3, padding=1, stride=1), dvgo.py:193-196)."""
    out = occ.copy()
    p = np.pad(occ, 1)
    n0, n1, n2 = occ.shape
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                out |= p[dx:dx + n0, dy:dy + n1, dz:dz + n2]
    return out


def act_shift_of(alpha_init):
    return float(F32(np.log(1.0 / (1.0 - alpha_init) - 1.0)))   # dvgo.py:49


def scene_params(reso, stepsize=0.5, alpha_init=1e-2, fast_color_thres=1e-4, bound=1.3):
    reso = (reso,) * 3 if isinstance(reso, int) else tuple(reso)
    voxel_size = (2 * bound) / reso[0]
    return dict(
        xyz_min=np.array([-bound] * 3, F32), xyz_max=np.array([bound] * 3, F32), reso=reso,
        near=2.0, far=1e9, stepsize=stepsize, stepdist=float(F32(stepsize * voxel_size)),
        voxel_size_ratio=1.0, interval=float(F32(stepsize * 1.0)), act_shift=act_shift_of(alpha_init),
        fast_color_thres=fast_color_thres, bg=1.0,
        weight_main=1.0, weight_entropy_last=0.001, weight_rgbper=0.01,
        lr_density=0.1, lr_k0=0.1, lr_net=1e-3, eps=1e-8, beta0=0.9, beta1=0.99, den_mode=1, k0_mode=1)


def _voxel_centers(reso, bound):
    ax = [np.linspace(-bound, bound, r, dtype=np.float64) for r in reso]
    return np.meshgrid(*ax, indexing="ij")


def mic_occupancy(reso, bound=1.3):
    """Union of a sphere (mic head), a capsule (stem) and a disc (base); ~1-3 % of the voxels."""
    X, Y, Z = _voxel_centers(reso, bound)
    sphere = (X ** 2 + Y ** 2 + (Z - 0.55) ** 2) <= 0.28 ** 2
    a, b = np.array([0, 0, 0.3]), np.array([0.35, 0, -0.75])
    ab = b - a
    t = ((X - a[0]) * ab[0] + (Y - a[1]) * ab[1] + (Z - a[2]) * ab[2]) / (ab @ ab)
    t = np.clip(t, 0, 1)
    d2 = (X - (a[0] + t * ab[0])) ** 2 + (Y - (a[1] + t * ab[1])) ** 2 + (Z - (a[2] + t * ab[2])) ** 2
    capsule = d2 <= 0.05 ** 2
    disc = (((X - 0.35) ** 2 + Y ** 2) <= 0.45 ** 2) & (np.abs(Z + 0.8) <= 0.03)
    return sphere | capsule | disc


def shell_occupancy(reso, bound=1.3, seed=6):
    """S512 stress scene: noisy thick shell around r = 0.8 (about 5 % of the voxels)."""
    X, Y, Z = _voxel_centers(reso, bound)
    r = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, (4, 3))
    noise = sum(0.5 ** k * np.sin((2 ** k) * 3.0 * X + ph[k, 0]) * np.sin((2 ** k) * 3.0 * Y + ph[k, 1]) *
                np.sin((2 ** k) * 3.0 * Z + ph[k, 2]) for k in range(4))
    return np.abs(r - 0.8 - 0.06 * noise) <= 0.022


def make_scene(reso=160, variant="dense", seed_density=1, seed_k0=2, seed_leaf=3, occupancy="mic", bound=1.3):
    """Returns a dict with `occ`, `active` (None = dense fill), `density` [R^3], `k0` [R^3,12], `mask` and params.

    variant "dense": denseFill topology (every 8^3 block is a leaf).  variant "sparse" (BASELINE config 2):
    each leaf block is dropped with p = 0.7 (seed 3); the tree holds the kept blocks only (all their voxels
    active) and the occupancy is restricted to them.
    """
    P = scene_params(reso, bound=bound)
    R = P["reso"]
    occ = mic_occupancy(R, bound) if occupancy == "mic" else shell_occupancy(R, bound)
    active = None
    if variant == "sparse":
        nb = tuple((r + 7) // 8 for r in R)
        keep = np.random.default_rng(seed_leaf).random(nb) >= 0.7
        # blocks that hold occupied voxels of a kept... keep is applied to every block alike
        active = np.repeat(np.repeat(np.repeat(keep, 8, 0), 8, 1), 8, 2)[:R[0], :R[1], :R[2]]
        occ = occ & active
    elif variant != "dense":
        raise ValueError(variant)
    dens = np.full(R, -10.0, F32)
    dens[occ] = np.random.default_rng(seed_density).normal(6.0, 1.0, int(occ.sum())).astype(F32)
    occ_d = _dilate(occ)
    k0 = np.zeros(R + (12,), F32)
    k0[occ_d] = np.random.default_rng(seed_k0).uniform(-1, 1, (int(occ_d.sum()), 12)).astype(F32)
    if active is not None:   # values outside the tree are background 0
        dens[~active] = 0.0
        k0[~active] = 0.0
    P.update(occ=occ, active=active, density=dens, k0=k0, mask=occ_d, variant=variant)
    return P


def mask_scale_shift(mask_shape, xyz_min, xyz_max):
    """MaskGrid buffers (plenvdb/lib/grid.py:229-231) in float32, like torch computes them."""
    xyz_min = np.asarray(xyz_min, F32)
    xyz_max = np.asarray(xyz_max, F32)
    scale = (np.asarray(mask_shape, F32) - F32(1)) / (xyz_max - xyz_min)
    shift = -xyz_min * scale
    return scale.astype(F32), shift.astype(F32)


# ---- cameras and rays
def pose_spherical(theta, phi, radius):
    """load_blender.py:29-34 (float32 matrices like torch.Tensor)."""
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], F32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], F32)

    def rot_theta(th):
        return np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], F32)

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180.0 * np.pi) @ c2w
    c2w = rot_theta(theta / 180.0 * np.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], F32) @ c2w
    return c2w.astype(F32)


def train_cameras(n=100, seed=4, radius=4.0):
    rng = np.random.default_rng(seed)
    th = rng.uniform(-180, 180, n)
    ph = rng.uniform(-90, 0, n)
    return np.stack([pose_spherical(t, p, radius) for t, p in zip(th, ph)])


def render_cameras(n=200, phi=-30.0, radius=4.0):
    return np.stack([pose_spherical(a, phi, radius) for a in np.linspace(-180, 180, n + 1)[:-1]])


def intrinsics(H=800, W=800, camera_angle_x=0.6911112070083618):
    focal = 0.5 * W / np.tan(0.5 * camera_angle_x)
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], F32)


def rays_of_pixels(K, c2w, px, py, inverse_y=False):
    """get_rays mode 'center' (dvgo.py:471-499) for pixel columns px, rows py (float32)."""
    i = px.astype(F32) + F32(0.5)
    j = py.astype(F32) + F32(0.5)
    if inverse_y:
        dirs = np.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], np.ones_like(i)], -1)
    else:
        dirs = np.stack([(i - K[0, 2]) / K[0, 0], -(j - K[1, 2]) / K[1, 1], -np.ones_like(i)], -1)
    dirs = dirs.astype(F32)
    rays_d = (dirs[..., None, :] * c2w[..., :3, :3]).sum(-1, dtype=F32).astype(F32)
    rays_o = np.broadcast_to(c2w[..., :3, 3], rays_d.shape).astype(F32)
    viewdirs = (rays_d / np.sqrt((rays_d * rays_d).sum(-1, keepdims=True, dtype=F32))).astype(F32)
    return np.ascontiguousarray(rays_o), np.ascontiguousarray(rays_d), np.ascontiguousarray(viewdirs)


def ray_batch(n_rays=8192, poses=None, K=None, H=800, W=800, seed=777, target_seed=5, inverse_y=False):
    """`n_rays` rays drawn uniformly from the (camera, pixel) pool, plus U(0,1) target colours."""
    poses = train_cameras() if poses is None else poses
    K = intrinsics(H, W) if K is None else K
    rng = np.random.default_rng(seed)
    cam = rng.integers(0, len(poses), n_rays)
    py = rng.integers(0, H, n_rays)
    px = rng.integers(0, W, n_rays)
    ro, rd, vd = rays_of_pixels(K, poses[cam], px, py, inverse_y)
    target = np.random.default_rng(target_seed).uniform(0, 1, (n_rays, 3)).astype(F32)
    return ro, rd, vd, target


def rgbnet_init(seed=777, k0_dim=12, pe_dim=27, width=128):
    """nn.Linear default init under torch.manual_seed (dvgo.py:99-107), packed w0,b0,w1,b1,w2,b2."""
    import torch
    g = torch.Generator().manual_seed(seed)

    def linear(fan_in, fan_out, zero_bias=False):
        bound = 1.0 / math.sqrt(fan_in)
        w = (torch.rand((fan_out, fan_in), generator=g) * 2 - 1) * bound   # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
        b = torch.zeros(fan_out) if zero_bias else (torch.rand(fan_out, generator=g) * 2 - 1) * bound
        return w, b
    w0, b0 = linear(k0_dim + pe_dim, width)
    w1, b1 = linear(width, width)
    w2, b2 = linear(width, 3, zero_bias=True)   # nn.init.constant_(rgbnet[-1].bias, 0)
    return torch.cat([t.reshape(-1) for t in (w0, b0, w1, b1, w2, b2)]).numpy().astype(F32)


def unpack_net(net, k0_dim=12, pe_dim=27, width=128):
    din = k0_dim + pe_dim
    o = 0
    out = []
    for shape in ((width, din), (width,), (width, width), (width,), (3, width), (3,)):
        n = int(np.prod(shape))
        out.append(net[o:o + n].reshape(shape))
        o += n
    return out
