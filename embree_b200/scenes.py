"""Synthetic scenes and ray sets of the benchmark / parity configurations.

Geometry generators restate the reference's own (tutorials/common/scenegraph/geometry_creation.cpp:8-35 plane,
:121-178 sphere; tutorials/triangle_geometry/triangle_geometry_device.cpp:31-97 cube + ground plane) so that the
"1 M" / "10 M" triangle scenes are exactly the verify benchmark's (tutorials/verify/verify.cpp:5757-6060:
createTriangleSphere(center 0, r 1, numPhi 501) = 1 002 000 triangles; numPhi 1581 = 9 991 920).
Ray generators follow the reference's definitions of "coherent" and "incoherent" and the path tracer's
diffuse bounce (tutorials/pathtracer/pathtracer_device.cpp:1119-1120,1597-1600; tutorials/common/math/sampling.h:52-78;
tutorials/common/math/random_sampler.h:15-80).  Rays are produced as torch tensors on any device so bench.py can
generate them directly in HBM; tests use the CPU.
"""
import math

import numpy as np
import torch

INVALID = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------------
# geometry (numpy, float32 / uint32)
# ------------------------------------------------------------------------------------------------------
def triangle_sphere(num_phi, center=(0.0, 0.0, 0.0), radius=1.0):
    """createTriangleSphere (geometry_creation.cpp:121-178): 2*numPhi meridians, numPhi+1 rings."""
    num_theta = 2 * num_phi
    f32 = np.float32
    rcp_t, rcp_p = f32(1.0) / f32(num_theta), f32(1.0) / f32(num_phi)
    phi = (np.arange(num_phi + 1, dtype=np.float32) * f32(math.pi) * rcp_p)[:, None]
    theta = (np.arange(num_theta, dtype=np.float32) * f32(2.0) * f32(math.pi) * rcp_t)[None, :]
    v = np.empty((num_phi + 1, num_theta, 3), np.float32)
    v[..., 0] = f32(center[0]) + f32(radius) * np.sin(phi) * np.sin(theta)
    v[..., 1] = f32(center[1]) + f32(radius) * np.cos(phi) * np.ones_like(theta)
    v[..., 2] = f32(center[2]) + f32(radius) * np.sin(phi) * np.cos(theta)
    th = np.arange(1, num_theta + 1, dtype=np.int64)
    tris = []
    # phi == 1 (note: the reference's p00 is the constant numTheta-1, kept as is)
    p10 = 1 * num_theta + th - 1
    p11 = 1 * num_theta + th % num_theta
    tris.append(np.stack([p10, np.full_like(th, num_theta - 1), p11], 1))
    for ph in range(2, num_phi):
        p00 = (ph - 1) * num_theta + th - 1
        p01 = (ph - 1) * num_theta + th % num_theta
        p10 = ph * num_theta + th - 1
        p11 = ph * num_theta + th % num_theta
        quad = np.stack([np.stack([p10, p00, p11], 1), np.stack([p01, p11, p00], 1)], 1).reshape(-1, 3)
        tris.append(quad)
    p00 = (num_phi - 1) * num_theta + th - 1
    p01 = (num_phi - 1) * num_theta + th % num_theta
    tris.append(np.stack([np.full_like(th, num_phi * num_theta), p00, p01], 1))
    return v.reshape(-1, 3), np.concatenate(tris, 0).astype(np.uint32)


def triangle_plane(p0, dx, dy, width, height):
    """createTrianglePlane (geometry_creation.cpp:8-35)."""
    p0, dx, dy = (np.asarray(a, np.float32) for a in (p0, dx, dy))
    xs = (np.arange(width + 1, dtype=np.float32) / np.float32(width))[None, :, None]
    ys = (np.arange(height + 1, dtype=np.float32) / np.float32(height))[:, None, None]
    v = (p0[None, None, :] + xs * dx[None, None, :] + ys * dy[None, None, :]).astype(np.float32)
    y, x = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    p00 = y * (width + 1) + x
    p01 = p00 + 1
    p10 = p00 + (width + 1)
    p11 = p10 + 1
    t = np.stack([np.stack([p00, p01, p10], -1), np.stack([p11, p10, p01], -1)], 2).reshape(-1, 3)
    return v.reshape(-1, 3), t.astype(np.uint32)


def terrain(n, seed=7, amplitude=0.15, octaves=5):
    """Non-convex companion scene (SURVEY 8d "S10b"): an n x n displaced plane over [-1,1]^2 in xz, height from
    seeded value noise.  2*n*n triangles (n = 2236 -> 9 999 392)."""
    v, t = triangle_plane((-1.0, 0.0, -1.0), (2.0, 0.0, 0.0), (0.0, 0.0, 2.0), n, n)
    rng = np.random.RandomState(seed)
    h = np.zeros((n + 1, n + 1), np.float32)
    u = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
    for o in range(octaves):
        cells = 4 << o
        g = rng.rand(cells + 2, cells + 2).astype(np.float32)
        fx = u * cells
        ix = np.minimum(fx.astype(np.int64), cells - 1)
        tx = fx - ix
        tx = tx * tx * (3.0 - 2.0 * tx)
        a = g[ix][:, ix] * (1 - tx)[None, :] + g[ix][:, ix + 1] * tx[None, :]
        b = g[ix + 1][:, ix] * (1 - tx)[None, :] + g[ix + 1][:, ix + 1] * tx[None, :]
        h += ((a * (1 - tx)[:, None] + b * tx[:, None]) - 0.5) * np.float32(amplitude / (1 << o))
    v = v.copy()
    v[:, 1] = h.reshape(-1)
    return v, t


def quad_terrain(n, seed=7, amplitude=0.15):
    """(n+1)^2 height-field vertices over [-1,1]^2, n*n NON-planar quads (v0,v1,v2,v3 counter-clockwise seen from +y)."""
    v, _ = terrain(n, seed=seed, amplitude=amplitude)
    i, j = np.meshgrid(np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32), indexing="ij")
    a = (i * (n + 1) + j).reshape(-1)
    q = np.stack([a, a + 1, a + n + 2, a + n + 1], axis=1).astype(np.uint32)
    return v, q


def hair_ball(strands, segments, seed=3, radius=1.0, length=0.35, width=0.004):
    """A fur ball in the spirit of tutorials/hair_geometry: `strands` curves of `segments` round linear segments each grow
    out of a sphere with a random walk and taper towards the tip.  Returns (vertices[nv,4] = xyz + radius, first-vertex
    index per segment, None) -- the neighbour flags are left to the library, as the tutorial does."""
    rng = np.random.RandomState(seed)
    n = rng.normal(size=(strands, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    step = length / segments
    pts = np.zeros((strands, segments + 1, 4), np.float32)
    p, d = n * radius, n.copy()
    for k in range(segments + 1):
        pts[:, k, :3] = p
        pts[:, k, 3] = width * (1.0 - 0.8 * k / segments)
        d = d + rng.normal(scale=0.35, size=d.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = p + d * step
    verts = pts.reshape(-1, 4)
    idx = (np.arange(strands)[:, None] * (segments + 1) + np.arange(segments)[None, :]).reshape(-1).astype(np.uint32)
    return verts, idx, None


def cubic_hair(strands, basis="bezier", knots=10, seed=7, radius=1.0, step=0.08, width=0.02):
    """Strands of connected cubic curves growing out of a sphere (tutorials/hair_geometry with --convert-...-to-curves, curve_geometry):
    `knots` float4 control points per strand from a random walk, tapering.  Index buffer = first control vertex of every
    curve: Bezier curves share their end points (0, 3, 6, ...), B-spline / Catmull-Rom curves slide by one, Hermite curves
    join consecutive vertices and come with a tangent per vertex (the strand's growth direction there, slightly perturbed, and the
    radius derivative).
    Returns (vertices[nv,4], indices[nc] u32, tangents[nv,4] or None)."""
    rng = np.random.RandomState(seed)
    n = rng.normal(size=(strands, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    pts = np.zeros((strands, knots, 4), np.float32)
    dirs = np.zeros((strands, knots, 3), np.float32)
    p, d = n * radius, n.copy()
    for k in range(knots):
        pts[:, k, :3] = p
        pts[:, k, 3] = width * (1.0 - 0.7 * k / knots)
        dirs[:, k] = d
        d = d + rng.normal(scale=0.5, size=d.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = p + d * step
    verts = pts.reshape(-1, 4)
    if basis == "bezier":
        starts = np.arange(0, knots - 3, 3)
    elif basis == "hermite":
        starts = np.arange(0, knots - 1)
    else:
        starts = np.arange(0, knots - 3)
    idx = (np.arange(strands)[:, None] * knots + starts[None, :]).reshape(-1).astype(np.uint32)
    tang = None
    if basis == "hermite":
        tang = np.zeros_like(verts)
        tang[:, :3] = dirs.reshape(-1, 3) * (1.1 * step) + rng.normal(scale=0.15 * step, size=(len(verts), 3))
        tang[:, 3] = -0.7 * width / knots + rng.normal(scale=0.02 * width, size=len(verts))
    return verts, idx, tang


def cube_and_ground():
    """The triangle_geometry tutorial scene (triangle_geometry_device.cpp:31-97): unit cube (12 tris, geomID 0)
    and a ground plane (2 tris, geomID 1)."""
    cv = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], np.float32)
    ct = np.array([[0, 1, 2], [1, 3, 2], [4, 6, 5], [5, 6, 7], [0, 4, 1], [1, 4, 5], [2, 3, 6], [3, 7, 6], [0, 2, 4], [2, 6, 4],
                   [1, 5, 3], [3, 5, 7]], np.uint32)
    gv = np.array([[-10, -2, -10], [-10, -2, 10], [10, -2, -10], [10, -2, 10]], np.float32)
    gt = np.array([[0, 1, 2], [1, 3, 2]], np.uint32)
    return (cv, ct), (gv, gt)


# ------------------------------------------------------------------------------------------------------
# random numbers of the tutorials (random_sampler.h:15-80), vectorised over ray ids
# ------------------------------------------------------------------------------------------------------
def _u32(x):
    return x & 0xFFFFFFFF


def _murmur_mix(h, k):
    k = _u32(k * 0xCC9E2D51)
    k = _u32((k << 15) | (k >> 17))
    k = _u32(k * 0x1B873593)
    h = h ^ k
    h = _u32((h << 13) | (h >> 19))
    return _u32(h * 5 + 0xE6546B64)


def _murmur_fin(h):
    h = h ^ (h >> 16)
    h = _u32(h * 0x85EBCA6B)
    h = h ^ (h >> 13)
    h = _u32(h * 0xC2B2AE35)
    return h ^ (h >> 16)


class RandomSampler:
    """State is an int64 tensor holding uint32 values (torch has no uint32 arithmetic)."""

    def __init__(self, ids):
        self.s = _murmur_fin(_murmur_mix(torch.zeros_like(ids), ids))

    def get1d(self):
        self.s = _u32(self.s * 1664525 + 1013904223)
        return (self.s >> 1).to(torch.float32) * 4.656612873077392578125e-10


# ------------------------------------------------------------------------------------------------------
# rays (torch).  A ray set is a float32 tensor [n, 24] viewed as RTCRayHit records (96 B): columns
# 0..2 org, 3 tnear, 4..6 dir, 7 time, 8 tfar, 9 mask, 10 id, 11 flags, 12..14 Ng, 15 u, 16 v, 17 primID, 18 geomID,
# 19 instID, 20 instPrimID, 21..23 padding.  Integer fields are written through an int32 view.
# ------------------------------------------------------------------------------------------------------
def pack_rayhits(org, dir, tnear, tfar, out=None):
    n = org.shape[0]
    r = out if out is not None else torch.empty((n, 24), dtype=torch.float32, device=org.device)
    r[:, 0:3] = org
    r[:, 3] = tnear
    r[:, 4:7] = dir
    r[:, 7] = 0.0
    r[:, 8] = tfar
    ri = r.view(torch.int32)
    ri[:, 9] = -1
    ri[:, 10] = torch.arange(n, device=org.device, dtype=torch.int32)
    ri[:, 11] = 0
    r[:, 12:17] = 0.0
    ri[:, 17:21] = -1
    ri[:, 21:24] = 0
    return r


def primary_rays(width, height, eye=(0.0, 0.0, 0.0), look=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), fov=90.0, device="cpu"):
    """Pinhole camera, dir = normalize(x*vx + y*vy + vz) as triangle_geometry_device.cpp:127; row-major pixels."""
    eye_t = torch.tensor(eye, dtype=torch.float32, device=device)
    w = torch.tensor(look, dtype=torch.float32, device=device)
    w = w / w.norm()
    upv = torch.tensor(up, dtype=torch.float32, device=device)
    u = torch.linalg.cross(upv, w)
    u = u / u.norm()
    v = torch.linalg.cross(w, u)
    fl = 0.5 * height / math.tan(0.5 * math.radians(fov))
    ys, xs = torch.meshgrid(torch.arange(height, device=device, dtype=torch.float32),
                            torch.arange(width, device=device, dtype=torch.float32), indexing="ij")
    d = (xs.reshape(-1, 1) - 0.5 * width + 0.5) * u + (0.5 * height - ys.reshape(-1, 1) - 0.5) * v + fl * w
    d = d / d.norm(dim=1, keepdim=True)
    org = eye_t.expand_as(d)
    return pack_rayhits(org, d, 0.0, float("inf"))


def verify_coherent_rays(width, height, tile=32, device="cpu"):
    """The reference's coherent benchmark rays (tutorials/verify/verify.cpp:5830-5897, MODE_INTERSECT16): origin 0,
    dir = (x/W, 1, y/H) (not normalised), tnear 0, tfar inf; returned in the benchmark's packet order -- `tile` x `tile`
    pixel tiles row-major, inside a tile 4x4-pixel packets row-major, lane = 4*dy + dx."""
    assert width % tile == 0 and height % tile == 0 and tile % 4 == 0
    ty, py, dy, tx, px, dx = torch.meshgrid(torch.arange(height // tile, device=device), torch.arange(tile // 4, device=device),
                                            torch.arange(4, device=device), torch.arange(width // tile, device=device),
                                            torch.arange(tile // 4, device=device), torch.arange(4, device=device), indexing="ij")
    # order: tileY, tileX, packetY, packetX, dy, dx
    perm = (0, 3, 1, 4, 2, 5)
    y = (ty * tile + py * 4 + dy).permute(*perm).reshape(-1).to(torch.float32)
    x = (tx * tile + px * 4 + dx).permute(*perm).reshape(-1).to(torch.float32)
    d = torch.stack([x * (1.0 / width), torch.ones_like(x), y * (1.0 / height)], 1)
    return pack_rayhits(torch.zeros_like(d), d, 0.0, float("inf"))


def tile_order_16(width, height):
    """Permutation that packs 4x4-pixel tiles into consecutive groups of 16 (RTCRayHit16 packets of config 2)."""
    assert width % 4 == 0 and height % 4 == 0
    idx = torch.arange(width * height).reshape(height // 4, 4, width // 4, 4).permute(0, 2, 1, 3).reshape(-1)
    return idx


def incoherent_rays_reference(n, org=(0.0, 0.0, 0.0), device="cpu", first_id=0):
    """The verify benchmark's incoherent set (rtcore_helpers.h:200-222): dir = 2*rand3 - 1 (not normalised)."""
    ids = torch.arange(first_id, first_id + n, device=device, dtype=torch.int64)
    rs = RandomSampler(ids)
    d = torch.stack([rs.get1d(), rs.get1d(), rs.get1d()], 1) * 2.0 - 1.0
    o = torch.tensor(org, dtype=torch.float32, device=device).expand_as(d)
    return pack_rayhits(o, d, 0.0, float("inf"))


def diffuse_bounce_rays(rayhits, seed=0, replicate=1, ids=None):
    """Cosine-weighted bounce off the hits in `rayhits` (traced RTCRayHit records), as the path tracer does:
    P = org + t*dir, N = normalize(Ng) facing the incoming ray, wi = frame(N) * cosineSampleHemisphere(u1,u2),
    eps = 32 * 1.19209e-7 * max(|P|, t), org' = P + eps*N, tnear = eps, tfar = inf.
    ids is None: only hit records emit, `replicate` rays each, random numbers seeded by the output index.
    ids given  : row k of `rayhits` (which must be a hit) emits exactly one ray seeded by ids[k]."""
    if ids is None:
        ri = rayhits.view(torch.int32)
        r = rayhits[ri[:, 18] != -1]
        if replicate > 1:
            r = r.repeat_interleave(replicate, dim=0)
        ids = torch.arange(r.shape[0], device=r.device, dtype=torch.int64)
    else:
        r = rayhits
    ids = ids + seed * 7919
    n = r.shape[0]
    dev = r.device
    org, d, t = r[:, 0:3], r[:, 4:7], r[:, 8:9]
    P = org + t * d
    Ng = r[:, 12:15]
    N = Ng / Ng.norm(dim=1, keepdim=True).clamp_min(1e-30)
    flip = (N * d).sum(1, keepdim=True) > 0
    N = torch.where(flip, -N, N)
    rs = RandomSampler(ids)
    u1, u2 = rs.get1d(), rs.get1d()
    phi = 2.0 * math.pi * u1
    ct, st = torch.sqrt(u2), torch.sqrt(1.0 - u2)
    lx, ly, lz = torch.cos(phi) * st, torch.sin(phi) * st, ct
    # frame(N) (linearspace3.h:117-124)
    zero = torch.zeros_like(N[:, 0])
    dx0 = torch.stack([zero, N[:, 2], -N[:, 1]], 1)
    dx1 = torch.stack([-N[:, 2], zero, N[:, 0]], 1)
    pick = ((dx0 * dx0).sum(1) > (dx1 * dx1).sum(1)).unsqueeze(1)
    dx = torch.where(pick, dx0, dx1)
    dx = dx / dx.norm(dim=1, keepdim=True)
    dy = torch.linalg.cross(N, dx)
    dy = dy / dy.norm(dim=1, keepdim=True)
    wi = lx.unsqueeze(1) * dx + ly.unsqueeze(1) * dy + lz.unsqueeze(1) * N
    wi = wi / wi.norm(dim=1, keepdim=True)
    eps = 32.0 * 1.19209e-7 * torch.maximum(P.abs().max(dim=1).values, t.squeeze(1))
    o2 = P + eps.unsqueeze(1) * N
    return pack_rayhits(o2, wi, eps, float("inf"))


def as_numpy_rayhits(t):
    """torch [n,24] float32 (CPU) -> numpy structured RTCRayHit[] sharing memory when possible."""
    from .rtc import RAYHIT_DTYPE, aligned_empty
    a = t.detach().cpu().contiguous().numpy()
    out = aligned_empty(a.shape[0], RAYHIT_DTYPE)
    out.view(np.float32).reshape(-1, 24)[:] = a
    return out


# ------------------------------------------------------------------------------------------------------
# path stream (BASELINE configs[4]): torch restatement of embree_b200/csrc/pathstream.cu -- the wavefront form of
# tutorials/pathtracer/pathtracer_device.cpp:1489-1603.  Used by the tests as the checker of the CUDA kernels and by
# bench.py's reference arm to generate the bounce streams on the CPU (untimed).
# ------------------------------------------------------------------------------------------------------
def camera_basis(width, height, eye, look, up=(0.0, 1.0, 0.0), fov=90.0):
    """cam12 = (p, U, V, W0) with dir(fx, fy) = normalize(fx*U + fy*V + W0): the pinhole of primary_rays()."""
    w = np.asarray(look, np.float64)
    w /= np.linalg.norm(w)
    u = np.cross(np.asarray(up, np.float64), w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    fl = 0.5 * height / math.tan(0.5 * math.radians(fov))
    w0 = fl * w - 0.5 * width * u + 0.5 * height * v
    return np.concatenate([np.asarray(eye, np.float64), u, -v, w0]).astype(np.float32)


def path_primary(first_path, n, cam12, width, height, spp, device="cpu"):
    """-> (rayhits [n,24], rng int64 [n], Lw [n]): jittered primary rays, pixel = path // spp, sample = path % spp."""
    path = torch.arange(first_path, first_path + n, device=device, dtype=torch.int64)
    pixel, sample = path // spp, path % spp
    x, y = pixel % width, pixel // width
    s = _murmur_fin(_murmur_mix(_murmur_mix(torch.zeros_like(path), x | (y << 16)), sample))

    def get1d():
        nonlocal s
        s = _u32(s * 1664525 + 1013904223)
        return (s >> 1).to(torch.float32) * 4.656612873077392578125e-10
    fx, fy = x.to(torch.float32) + get1d(), y.to(torch.float32) + get1d()
    get1d()
    c = torch.tensor(cam12, dtype=torch.float32, device=device)
    d = fx.unsqueeze(1) * c[3:6] + fy.unsqueeze(1) * c[6:9] + c[9:12]
    d = d / d.norm(dim=1, keepdim=True)
    r = pack_rayhits(c[0:3].expand_as(d), d, 0.0, float("inf"))
    return r, s, torch.ones(n, dtype=torch.float32, device=device)


def path_bounce(r, rng, Lw, light5):
    """One bounce on traced records `r` ([n,24], modified in place into the next rays).  Returns (shadow [n,12], rng, Lw,
    pending [n]).  light5 = (px, py, pz, intensity, albedo)."""
    n, dev = r.shape[0], r.device
    ri = r.view(torch.int32)
    alive = (r[:, 8] >= 0) & (ri[:, 18] != -1)
    org, d, t = r[:, 0:3].clone(), r[:, 4:7].clone(), r[:, 8:9].clone()
    P = org + t * d
    eps = 32.0 * 1.19209e-07 * torch.maximum(P.abs().max(dim=1).values, t.squeeze(1))
    Ng = r[:, 12:15] / r[:, 12:15].norm(dim=1, keepdim=True).clamp_min(1e-30)
    Ng = torch.where(((d * Ng).sum(1, keepdim=True) >= 0), -Ng, Ng)
    s = rng.clone()

    def get1d():
        nonlocal s
        s = _u32(s * 1664525 + 1013904223)
        return (s >> 1).to(torch.float32) * 4.656612873077392578125e-10
    u1, u2 = get1d(), get1d()
    phi, ct, st = 2.0 * math.pi * u1, torch.sqrt(u2), torch.sqrt(1.0 - u2)
    zero = torch.zeros_like(Ng[:, 0])
    dx0 = torch.stack([zero, -Ng[:, 2], Ng[:, 1]], 1)          # cross((1,0,0), N)
    dx1 = torch.stack([Ng[:, 2], zero, -Ng[:, 0]], 1)          # cross((0,1,0), N)
    dx = torch.where(((dx0 * dx0).sum(1) > (dx1 * dx1).sum(1)).unsqueeze(1), dx0, dx1)
    dx = dx / dx.norm(dim=1, keepdim=True).clamp_min(1e-30)
    dy = torch.linalg.cross(Ng, dx)
    dy = dy / dy.norm(dim=1, keepdim=True).clamp_min(1e-30)
    wi = (torch.cos(phi) * st).unsqueeze(1) * dx + (torch.sin(phi) * st).unsqueeze(1) * dy + ct.unsqueeze(1) * Ng
    get1d(); get1d()
    lp = torch.tensor(light5[0:3], dtype=torch.float32, device=dev)
    toL = lp - P
    dist2 = (toL * toL).sum(1)
    dist = torch.sqrt(dist2)
    ld = toL / dist.unsqueeze(1)
    cosl = (ld * Ng).sum(1).clamp_min(0.0)
    pending = torch.where(alive, Lw * (light5[3] / dist2) * (light5[4] * 0.318309886) * cosl, torch.zeros_like(Lw))
    shadow = torch.zeros((n, 12), dtype=torch.float32, device=dev)
    si = shadow.view(torch.int32)
    shadow[:, 0:3], shadow[:, 3], shadow[:, 4:7], shadow[:, 8] = P, eps, ld, dist
    si[:, 9], si[:, 10] = -1, ri[:, 10]
    dead = ~alive
    shadow[dead, 0:3] = 0.0
    shadow[dead, 3] = float("inf")
    shadow[dead, 4:7] = torch.tensor([0.0, 0.0, 1.0], device=dev)
    shadow[dead, 8] = float("-inf")
    Lw2 = torch.where(alive, Lw * light5[4], Lw)
    sign = torch.where((wi * Ng).sum(1) < 0, -1.0, 1.0)
    P2 = P + (sign * eps).unsqueeze(1) * Ng
    d2 = wi / wi.norm(dim=1, keepdim=True).clamp_min(1e-30)
    r[alive, 0:3], r[alive, 3], r[alive, 4:7], r[alive, 8] = P2[alive], eps[alive], d2[alive], float("inf")
    ri[alive, 17:20] = -1
    r[dead, 3] = float("inf")
    r[dead, 8] = float("-inf")
    return shadow, torch.where(alive, s, rng), Lw2, pending
