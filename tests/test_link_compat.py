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
