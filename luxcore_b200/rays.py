"""Ray-batch producers for tests and benchmarks.

Everything is written with torch ops so that the same code produces batches on the CPU (tests in
the build container) and directly in HBM on the GPU box (benchmarks: "rays must be generated on
the owning GPU").  A batch is a uint8 tensor [n, 48] holding packed luxrays::Ray records
(include/luxrays/core/geometry/ray.h:35-88); hits are uint8 [n, 20] luxrays::RayHit records.
"""
import math

import numpy as np
import torch

RAY_BYTES = 48
HIT_BYTES = 20
NULL_INDEX = 0xFFFFFFFF


def machine_epsilon(v, eps_min=1e-5, eps_max=1e-1):
    """MachineEpsilon::E (include/luxrays/core/epsilon.h:48-53,76-84): distance to the float 0x80
    ulps away, clamped to [min, max]."""
    v = v.to(torch.float32)
    nxt = (v.view(torch.int32) + 0x80).view(torch.float32)
    return torch.clamp((nxt - v).abs(), eps_min, eps_max)


def machine_epsilon_point(p):
    """E(Point) = max over components (epsilon.h:61-63)."""
    return machine_epsilon(p).amax(dim=-1)


def pack_rays(o, d, mint=None, maxt=None, time=None, flags=None):
    """-> uint8 [n, 48].  mint defaults to E(o), maxt to +inf (ray.h:38-46)."""
    n = o.shape[0]
    dev = o.device
    f = torch.zeros((n, 12), dtype=torch.float32, device=dev)
    f[:, 0:3] = o
    f[:, 3:6] = d
    f[:, 6] = machine_epsilon_point(o) if mint is None else mint
    f[:, 7] = float("inf") if maxt is None else maxt
    if time is not None:
        f[:, 8] = time
    if flags is not None:
        f[:, 9] = flags.to(torch.int32).view(torch.float32) if flags.dtype != torch.float32 else flags
    return f.view(torch.uint8).view(n, RAY_BYTES)


def rays_f32(rays_u8):
    return rays_u8.view(torch.float32).view(-1, 12)


def unpack_hits(hits_u8):
    """uint8 [n, 20] -> dict of tensors."""
    w = hits_u8.contiguous().view(torch.int32).view(-1, 5)
    f = w.view(torch.float32)
    return {"t": f[:, 0], "b1": f[:, 1], "b2": f[:, 2], "mesh": w[:, 3], "tri": w[:, 4]}


def _normalize(v):
    return v / v.norm(dim=-1, keepdim=True)


def camera_rays(cam, width, height, seed=1, device="cpu", jitter=True, time_range=None):
    """Pinhole camera rays, one stratified-jittered sample per pixel.
    cam = [orig(3), target(3), up(3), fov_degrees] (scene.camera.lookat / fieldofview)."""
    cam = [float(x) for x in cam]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    orig = torch.tensor(cam[0:3], dtype=torch.float32, device=device)
    target = torch.tensor(cam[3:6], dtype=torch.float32, device=device)
    up = torch.tensor(cam[6:9], dtype=torch.float32, device=device)
    fov = cam[9]
    w = _normalize(target - orig)
    u = _normalize(torch.linalg.cross(w, up))
    v = torch.linalg.cross(u, w)
    n = width * height
    idx = torch.arange(n, device=device)
    px = (idx % width).to(torch.float32)
    py = (idx // width).to(torch.float32)
    if jitter:
        j = torch.rand((n, 2), generator=g, device=device, dtype=torch.float32)
    else:
        j = torch.full((n, 2), 0.5, device=device, dtype=torch.float32)
    half = math.tan(math.radians(fov) * 0.5)
    aspect = width / float(height)
    sx = ((px + j[:, 0]) / width * 2.0 - 1.0) * half * (aspect if aspect > 1 else 1.0)
    sy = (1.0 - (py + j[:, 1]) / height * 2.0) * half * (1.0 / aspect if aspect < 1 else 1.0)
    d = _normalize(w[None, :] + sx[:, None] * u[None, :] + sy[:, None] * v[None, :])
    o = orig[None, :].expand(n, 3).contiguous()
    time = None
    if time_range is not None:
        time = time_range[0] + (time_range[1] - time_range[0]) * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    return pack_rays(o, d, time=time)


def uniform_rays(bbox_min, bbox_max, n, seed=3, device="cpu", time_range=None):
    """Stress batch: origins uniform in the scene box, directions uniform on the sphere."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lo = torch.as_tensor(bbox_min, dtype=torch.float32, device=device)
    hi = torch.as_tensor(bbox_max, dtype=torch.float32, device=device)
    o = lo + (hi - lo) * torch.rand((n, 3), generator=g, device=device, dtype=torch.float32)
    z = 1.0 - 2.0 * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    phi = 2.0 * math.pi * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    d = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1)
    time = None
    if time_range is not None:
        time = time_range[0] + (time_range[1] - time_range[0]) * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    return pack_rays(o, d, time=time)


def _concentric_disk(u1, u2):
    """ConcentricSampleDisk (src/luxrays/utils/mc.cpp:104-143), vectorised."""
    sx = 2.0 * u1 - 1.0
    sy = 2.0 * u2 - 1.0
    absx, absy = sx.abs(), sy.abs()
    use_x = absx > absy
    r = torch.where(use_x, sx, sy)
    safe = torch.where(use_x, sx, sy)
    safe = torch.where(safe == 0, torch.ones_like(safe), safe)
    theta = torch.where(use_x, (math.pi / 4.0) * (sy / safe), (math.pi / 2.0) - (math.pi / 4.0) * (sx / safe))
    return r * torch.cos(theta), r * torch.sin(theta)


def bounce_rays(rays_u8, hits_u8, tri_p0, tri_e1, tri_e2, seed=2):
    """Diffuse-bounce batch from a traced batch: for every hit,
    p = o + t d, n = normalised geometric normal flipped against d, d' = cosine-weighted hemisphere
    sample about n, origin pushed off the surface by E(p) along n, mint = E(o'), maxt = +inf.
    Misses are dropped.  tri_p0/e1/e2 are WORLD-space triangle vertex 0 and edges of the hit
    triangle, one row per ray (callers gather them with the hit's mesh/triangle index).
    Returns (rays uint8 [m, 48], index of the source ray [m])."""
    device = rays_u8.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    r = rays_f32(rays_u8)
    h = unpack_hits(hits_u8)
    hit = h["mesh"] != -1
    idx = torch.nonzero(hit, as_tuple=False)[:, 0]
    o = r[idx, 0:3]
    d = r[idx, 3:6]
    t = h["t"][idx]
    p = o + t[:, None] * d
    n = _normalize(torch.linalg.cross(tri_e1[idx], tri_e2[idx]))
    flip = (n * d).sum(dim=1) > 0
    n = torch.where(flip[:, None], -n, n)
    u = torch.rand((idx.shape[0], 2), generator=g, device=device, dtype=torch.float32)
    dx, dy = _concentric_disk(u[:, 0], u[:, 1])
    dz = torch.sqrt(torch.clamp(1.0 - dx * dx - dy * dy, min=0.0))
    # orthonormal frame about n
    a = torch.where((n[:, 0].abs() > 0.9)[:, None], torch.tensor([0.0, 1.0, 0.0], device=device).expand_as(n),
                    torch.tensor([1.0, 0.0, 0.0], device=device).expand_as(n))
    tx = _normalize(torch.linalg.cross(a, n))
    ty = torch.linalg.cross(n, tx)
    nd = _normalize(dx[:, None] * tx + dy[:, None] * ty + dz[:, None] * n)
    eps = machine_epsilon_point(p)
    no = p + n * eps[:, None]
    time = r[idx, 8]
    return pack_rays(no, nd, time=time), idx


def surface_rays(tri_p0, tri_e1, tri_e2, n, seed=7, device="cpu", axis_fraction=0.1):
    """Grazing stress batch: origins ON random triangles, moved along the geometric normal by
    k * E(p) with k uniform in [-2, 2] (exactly 0 for a quarter of them), directions uniform on the
    sphere; `axis_fraction` of the rays get an exactly axis-parallel direction (zero components ->
    infinite reciprocals).  These are the rays whose hits sit right at the planes of the boxes that
    bound the triangles they start from."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    p0 = torch.as_tensor(tri_p0, dtype=torch.float32, device=device)
    e1 = torch.as_tensor(tri_e1, dtype=torch.float32, device=device)
    e2 = torch.as_tensor(tri_e2, dtype=torch.float32, device=device)
    t = torch.randint(0, p0.shape[0], (n,), generator=g, device=device)
    u = torch.rand((n, 2), generator=g, device=device, dtype=torch.float32)
    su = torch.sqrt(u[:, 0])
    b1 = 1.0 - su
    b2 = u[:, 1] * su
    p = p0[t] + b1[:, None] * e1[t] + b2[:, None] * e2[t]
    nrm = torch.linalg.cross(e1[t], e2[t])
    ln = nrm.norm(dim=-1, keepdim=True)
    nrm = torch.where(ln > 0, nrm / torch.where(ln > 0, ln, torch.ones_like(ln)), torch.zeros_like(nrm))
    k = 4.0 * torch.rand(n, generator=g, device=device, dtype=torch.float32) - 2.0
    k = torch.where(torch.rand(n, generator=g, device=device) < 0.25, torch.zeros_like(k), k)
    o = p + nrm * (k * machine_epsilon_point(p))[:, None]
    z = 1.0 - 2.0 * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    phi = 2.0 * math.pi * torch.rand(n, generator=g, device=device, dtype=torch.float32)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    d = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1)
    ax = torch.rand(n, generator=g, device=device) < axis_fraction
    which = torch.randint(0, 6, (n,), generator=g, device=device)
    axes = torch.tensor([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=torch.float32, device=device)
    d = torch.where(ax[:, None], axes[which], d)
    return pack_rays(o, d)


def to_numpy_rays(rays_u8):
    from_dtype = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("mint", "<f4"), ("maxt", "<f4"), ("time", "<f4"),
                           ("flags", "<u4"), ("pad", "<f4", 2)])
    return rays_u8.cpu().numpy().reshape(-1).view(from_dtype)
