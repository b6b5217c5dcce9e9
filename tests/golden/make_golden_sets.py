"""Constants shared by make_golden.py (generator, needs the reference) and tests/conftest.py (loader, must not)."""
CUBIC_SETS = [("bezier", None), ("bspline", 7), ("catmull_rom", 4), ("hermite", 12)]   # (basis, tessellation rate; None = the default 4)


def point_scene(seed=17):
    """tutorials/point_geometry in small: three point sets around a triangle sphere (geomID 0) -- spheres (geomID 1, masked),
    ray-facing discs (geomID 2), oriented discs with random normals (geomID 3) -- and rays from around and inside the cloud."""
    import numpy as np
    from embree_b200 import scenes
    from embree_b200.rtc import make_rayhits
    rng = np.random.RandomState(seed)
    v, t = scenes.triangle_sphere(10)
    v = (v * np.float32(0.5)).astype(np.float32)

    def cloud(n, rmin, rmax):
        c = rng.normal(size=(n, 3)).astype(np.float32)
        c = c / np.linalg.norm(c, axis=1, keepdims=True) * rng.uniform(0.55, 1.3, (n, 1)).astype(np.float32)
        return np.concatenate([c, rng.uniform(rmin, rmax, (n, 1)).astype(np.float32)], 1).astype(np.float32)
    sets = [(cloud(700, 0.01, 0.09), "sphere", None, 1, 0x3), (cloud(500, 0.02, 0.1), "disc", None, 2, 0xFFFFFFFF),
            (cloud(500, 0.02, 0.12), "oriented_disc", rng.normal(size=(500, 3)).astype(np.float32), 3, 0xFFFFFFFD)]
    sets[0][0][5, 3] = -0.01          # negative radius: not a primitive (Points::valid)
    sets[1][0][7, 0] = np.nan         # invalid centre
    m = 8000
    org = rng.normal(size=(m, 3)).astype(np.float32)
    org = org / np.linalg.norm(org, axis=1, keepdims=True) * rng.uniform(0.6, 2.5, (m, 1)).astype(np.float32)
    k = np.arange(0, m, 9)
    org[k] = sets[0][0][rng.randint(8, 700, len(k)), :3] + np.float32(0.003)          # ray origins inside a sphere: back hits
    d = ((-org + rng.normal(scale=0.6, size=org.shape)) * rng.uniform(0.3, 3.0, (m, 1))).astype(np.float32)
    rays = make_rayhits(org, d)
    rays["tnear"][::7] = 0.05
    rays["tfar"][::5] = 0.8
    rays["mask"][::3] = 0x2
    rays["id"] = np.arange(len(rays))
    return [(v, t, 0, 0xFFFFFFFF)], sets, rays
