"""Filter-callback scenarios shared by tests/golden/make_golden.py (run on the unmodified reference, outputs committed as
tests/golden/filters.npz) and tests/test_filter.py (run on the B200 library through the same C-ABI and compared).

The callbacks are Python functions behind ctypes (rtc.FILTER_FUNCTION) that handle any N, as the reference's own test
callback does (tutorials/verify/verify.cpp:2771-2786): the reference calls them with N = 1 or the packet width, the B200
library always with N = 1.  The accept / reject rule is a pure function of the candidate's ids, so the final hit of a ray
does not depend on the order in which a traversal meets the candidates."""
import numpy as np

from embree_b200 import rtc
from embree_b200.rtc import FILTER_FUNCTION, RayQueryContext, filter_lane


def rejects(geomID, primID, instID=0xFFFFFFFF):
    h = (primID * 2654435761 + geomID * 40503 + (0 if instID == 0xFFFFFFFF else (instID + 1) * 97)) & 0xFFFFFFFF
    return ((h >> 7) & 3) == 0          # one candidate in four


class Recorder:
    """A filter callback + what it saw: (ray id, geomID, primID, instID, context instID, tfar, userPtr) per active lane."""

    def __init__(self, rule=rejects):
        self.calls = []
        self.rule = rule
        self.errors = []
        self.fn = FILTER_FUNCTION(self._call)

    def _call(self, args):
        try:
            a = args.contents
            for lane in range(a.N):
                if a.valid[lane] != -1:
                    continue
                f = filter_lane(args, lane)
                ctx_inst = a.context.contents.instID if a.context else None
                self.calls.append((f["id"], f["geomID"], f["primID"], f["instID"], ctx_inst, f["tfar"], a.geometryUserPtr))
                if self.rule(f["geomID"], f["primID"], f["instID"]):
                    a.valid[lane] = 0
        except Exception as e:   # an exception must not unwind through the C caller
            self.errors.append(repr(e))


def number_rays(rayhits):
    rayhits["id"] = np.arange(len(rayhits), dtype=np.uint32)
    return rayhits


def build(lib, dev, meshes, setup):
    """Scene of `meshes` [(v, t, geomID, mask)]; setup(lib, geomID, geometry handle) installs the callbacks."""
    sc = lib.rtcNewScene(dev)
    keep = []
    for (v, t, gid, mask) in meshes:
        _, k = lib.add_triangle_mesh(dev, sc, v, t, mask=mask, geom_id=gid)
        keep.append(k)
        setup(lib, gid, lib.rtcGetGeometry(sc, gid))
    lib.rtcCommitScene(sc)
    lib.check(dev)
    return sc, keep


CONFIGS = ["geometry", "argument_all", "argument_enabled"]


def run_config(lib, dev, meshes, rayhits, config, mode="1"):
    """One configuration on one library: returns (intersect_out, occluded_out, intersect Recorder, occluded Recorder).
      geometry         : geometry 0 carries an intersect and an occluded filter (rtcSetGeometry*FilterFunction), user data 123
      argument_all     : the arguments' filter with RTC_RAY_QUERY_FLAG_INVOKE_ARGUMENT_FILTER -> every geometry
      argument_enabled : the arguments' filter, enabled on the LAST geometry only (rtcSetGeometryEnableFilterFunctionFromArguments)"""
    ri, ro = Recorder(), Recorder()
    last = max(g for (_v, _t, g, _m) in meshes)

    def setup(L, gid, g):
        if config == "geometry" and gid == 0:
            L.rtcSetGeometryUserData(g, 123)
            L.rtcSetGeometryIntersectFilterFunction(g, ri.fn)
            L.rtcSetGeometryOccludedFilterFunction(g, ro.fn)
        if config == "argument_enabled" and gid == last:
            L.rtcSetGeometryEnableFilterFunctionFromArguments(g, True)
    sc, keep = build(lib, dev, meshes, setup)
    ctx = RayQueryContext(0xFFFFFFFF, 0xFFFFFFFF)
    if config == "geometry":
        ai = ao = lib.args(context=ctx)
    else:
        inv = config == "argument_all"
        ai = lib.args(filter=ri.fn, invoke_argument_filter=inv, context=ctx)
        ao = lib.args(filter=ro.fn, invoke_argument_filter=inv, context=ctx)
    out_i = lib.intersect(sc, number_rays(rayhits.copy()), mode, args=ai)
    out_o = lib.occluded(sc, rtc.rays_of(number_rays(rayhits.copy())), mode, args=ao)
    lib.check(dev)
    lib.rtcReleaseScene(sc)
    assert not ri.errors and not ro.errors, (ri.errors, ro.errors)
    return out_i, out_o, ri, ro
