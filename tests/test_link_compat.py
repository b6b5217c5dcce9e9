"""The reference's own tutorial (tutorials/minimal/minimal.cpp), compiled unmodified against the reference's own
headers, linked against libembree4_b200.so (tests/link_compat/build.sh): existing callers link unchanged."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "link_compat", "_bin", "embree_minimal")
LIB = os.path.join(ROOT, "embree_b200", "csrc", "libembree4_b200.so")


def _ensure_built():
    subprocess.check_call([os.path.join(ROOT, "tests", "link_compat", "build.sh")])
    if not os.path.exists(EXE):
        pytest.skip("reference sources not available and no prebuilt binary")


def test_reference_tutorial_links_against_our_library():
    _ensure_built()
    needed = subprocess.check_output(["readelf", "-d", EXE]).decode()
    assert "libembree4_b200.so" in needed and "libembree4.so.4" not in needed
    und = subprocess.check_output(["nm", "-D", "--undefined-only", EXE]).decode()
    wanted = sorted(l.split()[-1] for l in und.splitlines() if l.split()[-1].startswith("rtc"))
    assert "rtcIntersect1" in wanted and "rtcCommitScene" in wanted
    exported = subprocess.check_output(["nm", "-D", "--defined-only", LIB]).decode()
    have = set(l.split()[-1] for l in exported.splitlines())
    assert not [w for w in wanted if w not in have]


@pytest.mark.gpu
def test_reference_tutorial_runs_on_the_gpu():
    """tutorials/minimal/minimal.cpp:159-206 prints one hit (geometry 0, primitive 0, tfar=1) and one miss."""
    _ensure_built()
    out = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120).stdout
    assert "Found intersection on geometry 0, primitive 0 at tfar=1.000000" in out, out
    assert "Did not find any intersection." in out, out


TUT = os.path.join(ROOT, "tests", "link_compat", "_bin", "embree_triangle_geometry")
GOLDEN_FRAME = os.path.join(ROOT, "tests", "golden", "triangle_geometry_160x120.raw")


@pytest.mark.gpu
def test_triangle_geometry_tutorial_renders_the_reference_frame(tmp_path):
    """BASELINE configs[0]: the reference's tutorials/triangle_geometry device code (its own multi-threaded tile loop of
    rtcTraversableIntersect1 + rtcTraversableOccluded1 per pixel), compiled untouched and linked against libembree4_b200.so,
    renders the frame the same code produces with the unmodified reference library.  Pixels may differ only ON an edge of
    the picture (cube edges, silhouettes, the shadow boundary): there a ray hits two faces at the same distance or a shadow
    ray grazes, and the winner is order dependent in the reference itself; at most 0.3 % of the frame."""
    import numpy as np
    _ensure_built()
    if not os.path.exists(TUT):
        pytest.skip("tutorial binary not built")
    out = str(tmp_path / "frame.raw")
    r = subprocess.run([TUT, out, "160", "120", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "device error 0" in r.stdout, r.stdout
    got = np.fromfile(out, np.int32)
    want = np.fromfile(GOLDEN_FRAME, np.int32)
    assert got.shape == want.shape and len(np.unique(want)) >= 5
    g, w = got.reshape(120, 160), want.reshape(120, 160)
    edge = np.zeros_like(w, bool)                      # pixels of the golden frame with a differently coloured 8-neighbour
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            sh = np.roll(np.roll(w, dy, 0), dx, 1)
            edge |= sh != w
    diff = g != w
    assert diff.sum() <= 0.003 * w.size and not (diff & ~edge).any(), f"{diff.sum()} pixels differ, {(diff & ~edge).sum()} of them off an edge"


POINTS = os.path.join(ROOT, "tests", "link_compat", "_bin", "embree_point_geometry")


@pytest.mark.gpu
def test_point_geometry_tutorial_renders_the_reference_frame(tmp_path):
    """The reference's tutorials/point_geometry device code (512 sphere points, 512 ray-facing discs and 512 oriented discs over a ground
    plane; closest hit + shadow ray per pixel), compiled untouched and linked against libembree4_b200.so, renders the frame the same
    code produces with the unmodified reference library; pixels may differ only on an edge of the picture (a silhouette or a shadow
    boundary, where the reference's own approximate reciprocal decides), at most 0.5 % of the frame; one step of one 8-bit colour channel
    (the quantised shading of a normal that differs in its last bits) is tolerated anywhere, on at most 3 % of the frame."""
    import numpy as np
    _ensure_built()
    if not os.path.exists(POINTS):
        pytest.skip("tutorial binary not built")
    out = str(tmp_path / "frame.raw")
    r = subprocess.run([POINTS, out, "160", "120", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "device error 0" in r.stdout, r.stdout
    got = np.fromfile(out, np.int32)
    want = np.fromfile(os.path.join(ROOT, "tests", "golden", "point_geometry_160x120.raw"), np.int32)
    assert got.shape == want.shape and len(np.unique(want)) >= 100
    g, w = got.reshape(120, 160), want.reshape(120, 160)
    edge = np.zeros_like(w, bool)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            edge |= np.roll(np.roll(w, dy, 0), dx, 1) != w
    ch = lambda a: np.stack([(a >> sh) & 0xFF for sh in (0, 8, 16)], -1).astype(np.int32)      # noqa: E731
    big = (np.abs(ch(g) - ch(w)) > 1).any(-1)      # more than one step of one 8-bit channel (the shading quantises dot(light, normalize(Ng)))
    diff = g != w
    assert big.sum() <= 0.005 * w.size and diff.sum() <= 0.03 * w.size and not (big & ~edge).any(), f"{diff.sum()} pixels differ, {big.sum()} by more than one step, {(big & ~edge).sum()} of them off an edge"


HAIR = os.path.join(ROOT, "tests", "link_compat", "_bin", "embree_hair_geometry")


@pytest.mark.gpu
@pytest.mark.parametrize("ctype,name", [(0, "round linear"), (1, "flat Bezier"), (2, "round Bezier")])
def test_hair_geometry_tutorial_renders_the_reference_frame(tmp_path, ctype, name):
    """BASELINE configs[3]a: the reference's tutorials/hair_geometry device code -- hair sets converted with
    rtcSetGeometryTessellationRate / rtcSetGeometryEnableFilterFunctionFromArguments, per pixel a path of up to 21
    rtcTraversableIntersect1 calls each followed by an rtcTraversableOccluded1 shadow ray whose transparency-accumulating occlusion
    filter is passed in the arguments -- compiled untouched and linked against libembree4_b200.so, renders a procedural fur ball
    (tests/link_compat/hair_host.cpp) like the same code does with the unmodified reference library (golden frames).
    Flat Bezier hair: the frame is pixel-identical (the shading frame of a ribbon hit does not depend on where across the ribbon the
    ray lands).  Round hair (linear segments as in the shipped furBall model, and Bezier): a path bounces off tubes of radius
    0.002-0.006 at distances of ~0.1, which amplifies any difference of a hit ~25x per bounce, so paths of several bounces decorrelate:
    the reference's own AVX2 and AVX-512 code paths (hits differing by ~1e-7) already disagree on 0.5 % / 0.3 % of these pixels, this
    library (round-curve distances within ~1e-6..1e-5 of the reference's) on 4.3 % / 3.3 %.  There the frames must agree statistically:
    >= 93 % of the pixels identical and the mean colour within one level (seen: 0.17 levels)."""
    import numpy as np
    _ensure_built()
    if not os.path.exists(HAIR):
        pytest.skip("tutorial binary not built")
    out = str(tmp_path / "frame.raw")
    r = subprocess.run([HAIR, out, "96", "72", "4", str(ctype), "3000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "device error 0" in r.stdout, r.stdout
    got = np.fromfile(out, np.uint32)
    want = np.fromfile(os.path.join(ROOT, "tests", "golden", f"hair_geometry_{ctype}_96x72.raw"), np.uint32)
    assert got.shape == want.shape and len(np.unique(want)) >= 300

    def ch(a):
        return np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.int32)
    same = (got == want).mean()
    mean_diff = np.abs(ch(got).mean(0) - ch(want).mean(0)).max()
    if ctype == 1:
        assert same >= 0.995, (name, same)
    else:
        assert same >= 0.93 and mean_diff < 1.0, (name, same, mean_diff)


DYN = os.path.join(ROOT, "tests", "link_compat", "_bin", "embree_dynamic_scene")


@pytest.mark.gpu
def test_dynamic_scene_tutorial_renders_the_reference_frames(tmp_path):
    """BASELINE configs[3]b: the reference's tutorials/dynamic_scene device code -- an RTC_SCENE_FLAG_DYNAMIC | RTC_SCENE_FLAG_ROBUST scene of a
    plane and 20 spheres with per-geometry build qualities whose vertices it rewrites (rtcGetGeometryBufferData, rtcUpdateGeometryBuffer,
    rtcCommitGeometry) and re-commits every frame -- compiled untouched and linked against libembree4_b200.so, renders three frames of the
    animation like the same code does with the unmodified reference library (every sphere moves every frame, so each commit rebuilds one
    BVH over everything -- the library's own choice between that and the two-level path).  Pixels may differ only ON an edge of the picture (silhouettes, shadow boundaries), at most 0.5 % of a frame."""
    import numpy as np
    _ensure_built()
    if not os.path.exists(DYN):
        pytest.skip("tutorial binary not built")
    r = subprocess.run([DYN, str(tmp_path / "frame"), "160", "120", "4", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "device error 0" in r.stdout, r.stdout
    frames = []
    for f in range(3):
        got = np.fromfile(str(tmp_path / f"frame_{f}.raw"), np.int32)
        want = np.fromfile(os.path.join(ROOT, "tests", "golden", f"dynamic_scene_160x120_{f}.raw"), np.int32)
        assert got.shape == want.shape and len(np.unique(want)) >= 20
        g, w = got.reshape(120, 160), want.reshape(120, 160)
        edge = np.zeros_like(w, bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                edge |= np.roll(np.roll(w, dy, 0), dx, 1) != w
        diff = g != w
        assert diff.sum() <= 0.005 * w.size and not (diff & ~edge).any(), f"frame {f}: {diff.sum()} pixels differ, {(diff & ~edge).sum()} of them off an edge"
        frames.append(w)
    assert (frames[0] != frames[1]).mean() > 0.05 and (frames[1] != frames[2]).mean() > 0.05      # the scene really moves
