// trace.cu -- closest-hit / any-hit traversal kernels for sm_100a.
//
// Replaces kernels/bvh/bvh_intersector1.cpp:31-197 (single ray) and, through the packet I/O adapters below,
// kernels/bvh/bvh_intersector_hybrid.cpp:106-370,600-783 (rtcIntersect4/8/16, rtcOccluded4/8/16): on this
// device a "packet" is only an I/O layout -- every lane is traced as an independent ray of a 32-wide warp, which is
// what the reference's hybrid traverser degenerates to below its switch threshold (bvh_intersector_hybrid.h:33-37).
#include <stdio.h>

#include "rtk_device.h"

namespace rtk {

// ---- 16-byte loads of the BVH through the read-only path ------------------------------------------------------
struct NodeLoadG {
  const uint4* __restrict__ base;
  __device__ __forceinline__ u32x4 operator()(uint32_t node, int k) const {
    const uint4 v = __ldg(base + (size_t)node * 5 + k);
    return u32x4{v.x, v.y, v.z, v.w};
  }
};
struct TriLoadG {
  const uint4* __restrict__ base;
  __device__ __forceinline__ u32x4 operator()(uint32_t tri, int k) const {
    const uint4 v = __ldg(base + (size_t)tri * 3 + k);
    return u32x4{v.x, v.y, v.z, v.w};
  }
};

// ---- ray / hit I/O adapters ------------------------------------------------------------------------------------
// K == 1 : AoS RTCRayHit (96 B, hit at +48) or RTCRay (48 B).  K in {4,8,16}: SoA inside each packet
// (include/embree4/rtcore_ray.h:55-184): field f of lane l of packet m lives at m*PACKET + (f*K + l)*4.
template <int K, bool OCCLUDED>
struct RayIO;

template <bool OCCLUDED>
struct RayIO<1, OCCLUDED> {
  static constexpr int kStride = OCCLUDED ? 48 : 96;
  static __device__ __forceinline__ bool load(const TraceParams& p, unsigned long long i, Ray& r) {
    const float4* src = reinterpret_cast<const float4*>(static_cast<const char*>(p.rays) + i * kStride);
    const float4 a = src[0], b = src[1], c = src[2];
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tnear = a.w;
    r.dx = b.x; r.dy = b.y; r.dz = b.z; r.time = b.w;
    r.tfar = c.x; r.mask = __float_as_uint(c.y); r.id = __float_as_uint(c.z); r.flags = __float_as_uint(c.w);
    return true;
  }
  static __device__ __forceinline__ void store_tfar(const TraceParams& p, unsigned long long i, float tfar) {
    *reinterpret_cast<float*>(static_cast<char*>(p.rays) + i * kStride + 32) = tfar;
  }
  static __device__ __forceinline__ void store_hit(const TraceParams& p, unsigned long long i, const Hit& h) {
    char* rec = static_cast<char*>(p.rays) + i * kStride;
    *reinterpret_cast<float*>(rec + 32) = h.t;
    float4 a, b;
    a.x = h.ngx; a.y = h.ngy; a.z = h.ngz; a.w = h.u;
    b.x = h.v; b.y = __uint_as_float(h.primID); b.z = __uint_as_float(h.geomID); b.w = __uint_as_float(p.instID);
    *reinterpret_cast<float4*>(rec + 48) = a;
    *reinterpret_cast<float4*>(rec + 64) = b;
    *reinterpret_cast<uint32_t*>(rec + 80) = p.instPrimID;
  }
};

template <int K, bool OCCLUDED>
struct RayIO {
  static constexpr int kPacket = (OCCLUDED ? 12 : 21) * 4 * K;
  static __device__ __forceinline__ char* field(const TraceParams& p, unsigned long long i, int f) {
    return static_cast<char*>(p.rays) + (i / K) * kPacket + ((size_t)f * K + (i % K)) * 4;
  }
  static __device__ __forceinline__ bool load(const TraceParams& p, unsigned long long i, Ray& r) {
    if (p.valid && p.valid[i] != -1) return false;  // inactive lane: record must come back untouched
    r.ox = *reinterpret_cast<float*>(field(p, i, 0)); r.oy = *reinterpret_cast<float*>(field(p, i, 1));
    r.oz = *reinterpret_cast<float*>(field(p, i, 2)); r.tnear = *reinterpret_cast<float*>(field(p, i, 3));
    r.dx = *reinterpret_cast<float*>(field(p, i, 4)); r.dy = *reinterpret_cast<float*>(field(p, i, 5));
    r.dz = *reinterpret_cast<float*>(field(p, i, 6)); r.time = *reinterpret_cast<float*>(field(p, i, 7));
    r.tfar = *reinterpret_cast<float*>(field(p, i, 8)); r.mask = *reinterpret_cast<uint32_t*>(field(p, i, 9));
    r.id = 0; r.flags = 0;
    return true;
  }
  static __device__ __forceinline__ void store_tfar(const TraceParams& p, unsigned long long i, float tfar) {
    *reinterpret_cast<float*>(field(p, i, 8)) = tfar;
  }
  static __device__ __forceinline__ void store_hit(const TraceParams& p, unsigned long long i, const Hit& h) {
    *reinterpret_cast<float*>(field(p, i, 8)) = h.t;
    *reinterpret_cast<float*>(field(p, i, 12)) = h.ngx; *reinterpret_cast<float*>(field(p, i, 13)) = h.ngy;
    *reinterpret_cast<float*>(field(p, i, 14)) = h.ngz; *reinterpret_cast<float*>(field(p, i, 15)) = h.u;
    *reinterpret_cast<float*>(field(p, i, 16)) = h.v; *reinterpret_cast<uint32_t*>(field(p, i, 17)) = h.primID;
    *reinterpret_cast<uint32_t*>(field(p, i, 18)) = h.geomID; *reinterpret_cast<uint32_t*>(field(p, i, 19)) = p.instID;
    *reinterpret_cast<uint32_t*>(field(p, i, 20)) = p.instPrimID;
  }
};

constexpr int TRACE_THREADS = 128;
constexpr int TRACE_WARPS = TRACE_THREADS / 32;
constexpr int kTriWaitMax = 3;    // ... but never make a lane wait longer than this many iterations
constexpr int kTriBatchMin = 6;   // run a triangle step only when this many lanes wait for one (or no lane has node work)

// Persistent warps.  Every warp owns the 32-ray blocks w, w+W, w+2W, ... of the stream (W = warps in the grid) and
// keeps its 32 lanes busy: a lane whose ray has terminated writes its result and immediately takes the warp's next
// unassigned ray (ballot + popc ranking, no atomics).  One loop iteration = at most one node step (pop a child of the
// current node group, fetch the 80-byte node, slab-test its 8 children) for the lanes that want it, and at most one
// triangle step for the lanes that have triangle hits pending; the two phases are warp-synchronous so lanes in the
// same phase execute together instead of serialising through a per-thread while-while loop.
template <int K, bool OCCLUDED, bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS) trace_kernel(const TraceParams p) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned long long warp_id = (unsigned long long)blockIdx.x * TRACE_WARPS + (threadIdx.x >> 5);
  const unsigned long long num_warps = (unsigned long long)gridDim.x * TRACE_WARPS;
  const uint4* __restrict__ nodes = reinterpret_cast<const uint4*>(p.nodes);
  const uint4* __restrict__ tris = reinterpret_cast<const uint4*>(p.tris);

  // per-lane ray state
  Ray r;
  float idx = 0, idy = 0, idz = 0, tnear_c = 0, tfar_c = 0, tfar_tri = 0;
  uint32_t oct = 0, oct_inv4 = 0;
  bool negx = false, negy = false, negz = false;
  Hit hit;
  hit.t = 0; hit.u = 0; hit.v = 0; hit.ngx = 0; hit.ngy = 0; hit.ngz = 0; hit.primID = 0; hit.geomID = 0;
  bool found = false, active = false;
  unsigned long long ray_index = 0;
  uint32_t ngx = 0, ngy = 0, tgx = 0, tgy = 0;
  uint32_t stack_x[kStackSize], stack_y[kStackSize];
  int sp = 0;
  unsigned long long j_next = 0;        // warp-uniform: next unassigned ray of this warp's private sequence
  unsigned long long st_rays = 0, st_nodes = 0, st_tris = 0;
  bool exhausted = false;                // this lane was handed an index beyond the stream
  int tri_wait = 0;                      // warp-uniform: iterations some lane has been waiting for a triangle step

  for (;;) {
    // ---- 1. refill idle lanes
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle) {
      if (!active && !exhausted) {
        const unsigned long long j = j_next + __popc(idle & lt_mask);
        ray_index = ((j >> 5) * num_warps + warp_id) * 32ull + (j & 31ull);
        if (ray_index >= p.n) exhausted = true;
        else if (RayIO<K, OCCLUDED>::load(p, ray_index, r)) {
          if (STATS) ++st_rays;
          found = false;
          sp = 0; tgx = 0; tgy = 0;
          // empty scene / already occluded rays terminate at once (bvh_intersector1.cpp:39,128-129)
          const bool go = p.root_valid && !(OCCLUDED && r.tfar < 0.0f);
          idx = rcp_safe(r.dx); idy = rcp_safe(r.dy); idz = rcp_safe(r.dz);
          negx = idx < 0.0f; negy = idy < 0.0f; negz = idz < 0.0f;
          oct = (negx ? 1u : 0u) | (negy ? 2u : 0u) | (negz ? 4u : 0u);
          oct_inv4 = (7u - oct) * 0x01010101u;
          tnear_c = fmaxf(r.tnear, 0.0f);
          tfar_c = fmaxf(r.tfar, 0.0f);
          tfar_tri = r.tfar;
          ngx = 0; ngy = go ? 0x80000000u : 0u;   // root entered as "one pending internal child, imask 0"
          active = go;
        }
      }
      j_next += __popc(idle);
      if (!__any_sync(FULL, active)) {
        if (__all_sync(FULL, exhausted)) break;
        continue;
      }
    }
    // ---- 2. node step for lanes that have no triangle pending and a node child pending
    const bool want_node = active && tgy == 0 && (ngy & 0xFF000000u);
    if (want_node) {
      const int bit = 31 - __clz((int)ngy);
      ngy &= ~(1u << bit);
      if (ngy & 0xFF000000u) { stack_x[sp] = ngx; stack_y[sp] = ngy; ++sp; }
      const uint32_t slot = ((uint32_t)(bit - 24)) ^ (7u - oct);
      const uint32_t node_index = ngx + (uint32_t)__popc(ngy & 0xFFu & ((1u << slot) - 1u));
      const uint4* np = nodes + (size_t)node_index * 5;
      const uint4 a0 = __ldg(np), a1 = __ldg(np + 1), a2 = __ldg(np + 2), a3 = __ldg(np + 3), a4 = __ldg(np + 4);
      if (STATS) ++st_nodes;
      const u32x4 n0{a0.x, a0.y, a0.z, a0.w}, n1{a1.x, a1.y, a1.z, a1.w}, n2{a2.x, a2.y, a2.z, a2.w},
          n3{a3.x, a3.y, a3.z, a3.w}, n4{a4.x, a4.y, a4.z, a4.w};
      const uint32_t hm = node_hitmask<OCCLUDED>(n0, n1, n2, n3, n4, r.ox, r.oy, r.oz, idx, idy, idz, negx, negy, negz,
                                                 tnear_c, tfar_c, oct_inv4);
      ngx = n1.x;
      ngy = (hm & 0xFF000000u) | (n0.w >> 24);
      tgx = n1.y;
      tgy = hm & 0x00FFFFFFu;
    }
    // ---- 3. triangle step, batched across the warp
    const unsigned tri_lanes = __ballot_sync(FULL, active && tgy != 0);
    const unsigned node_lanes = __ballot_sync(FULL, active && tgy == 0 && (ngy & 0xFF000000u));
    if (tri_lanes && (__popc(tri_lanes) >= kTriBatchMin || node_lanes == 0 || ++tri_wait >= kTriWaitMax)) {
      tri_wait = 0;
      if (active && tgy != 0) {
        const int tb = 31 - __clz((int)tgy);
        tgy &= ~(1u << tb);
        const uint4* tp = tris + (size_t)(tgx + (uint32_t)tb) * 3;
        const uint4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (STATS) ++st_tris;
        TriHit th;
        if (tri_test(r, tfar_tri, __uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(b.x),
                     __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(c.x), __uint_as_float(c.y),
                     __uint_as_float(c.z), th) &&
            (c.w & r.mask) != 0) {
          found = true;
          if (OCCLUDED) { ngy = 0; tgy = 0; sp = 0; }     // any hit terminates the ray
          else {
            const float rcpAbsDen = 1.0f / th.absDen;
            hit.t = th.T * rcpAbsDen; hit.u = th.U * rcpAbsDen; hit.v = th.V * rcpAbsDen;
            hit.ngx = th.ngx; hit.ngy = th.ngy; hit.ngz = th.ngz;
            hit.primID = a.w; hit.geomID = b.w;
            tfar_tri = hit.t;
            tfar_c = fmaxf(hit.t, 0.0f);
          }
        }
      }
    }
    // ---- 4. pop or finish
    if (active && tgy == 0 && (ngy & 0xFF000000u) == 0) {
      if (sp > 0) {
        --sp;
        const uint32_t px = stack_x[sp], py = stack_y[sp];
        if (py & 0xFF000000u) { ngx = px; ngy = py; }
        else { tgx = px; tgy = py; ngx = 0; ngy = 0; }
      } else {
        if (found) {
          if (OCCLUDED) RayIO<K, OCCLUDED>::store_tfar(p, ray_index, -INFINITY);   // bvh_intersector1.cpp:186-188
          else RayIO<K, OCCLUDED>::store_hit(p, ray_index, hit);
        }
        active = false;
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      st_rays += __shfl_xor_sync(FULL, st_rays, o);
      st_nodes += __shfl_xor_sync(FULL, st_nodes, o);
      st_tris += __shfl_xor_sync(FULL, st_tris, o);
    }
    if (lane == 0) { atomicAdd(&p.stat[0], st_rays); atomicAdd(&p.stat[1], st_nodes); atomicAdd(&p.stat[2], st_tris); }
  }
}

static int g_num_sms = 0;

template <int K, bool OCCLUDED>
static int launch_k(const TraceParams& p, cudaStream_t st) {
  if (p.n == 0) return 0;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  // persistent grid: a multiple of the SM count, capped by the work available
  const unsigned long long need = (p.n + TRACE_THREADS - 1) / TRACE_THREADS;
  const unsigned long long cap = (unsigned long long)g_num_sms * 6;
  const unsigned blocks = (unsigned)(need < cap ? need : cap);
  if (p.stat) trace_kernel<K, OCCLUDED, true><<<blocks, TRACE_THREADS, 0, st>>>(p);
  else trace_kernel<K, OCCLUDED, false><<<blocks, TRACE_THREADS, 0, st>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_trace(const TraceParams& p, int occluded, int K, cudaStream_t st) {
  switch (K) {
    case 1: return occluded ? launch_k<1, true>(p, st) : launch_k<1, false>(p, st);
    case 4: return occluded ? launch_k<4, true>(p, st) : launch_k<4, false>(p, st);
    case 8: return occluded ? launch_k<8, true>(p, st) : launch_k<8, false>(p, st);
    case 16: return occluded ? launch_k<16, true>(p, st) : launch_k<16, false>(p, st);
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace rtk
