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

template <int K, bool OCCLUDED, bool STATS>
__global__ void __launch_bounds__(TRACE_THREADS) trace_kernel(const TraceParams p) {
  const unsigned long long i = (unsigned long long)blockIdx.x * TRACE_THREADS + threadIdx.x;
  TravStats st{0, 0};
  bool active = false;
  if (i < p.n) {
    Ray r;
    active = RayIO<K, OCCLUDED>::load(p, i, r);
    if (active) {
      Hit h;
      const NodeLoadG ldn{reinterpret_cast<const uint4*>(p.nodes)};
      const TriLoadG ldt{reinterpret_cast<const uint4*>(p.tris)};
      const bool found = traverse<OCCLUDED, STATS>(r, h, ldn, ldt, p.root_valid, &st);
      if (found) {
        if (OCCLUDED) RayIO<K, OCCLUDED>::store_tfar(p, i, -INFINITY);  // bvh_intersector1.cpp:186-188
        else RayIO<K, OCCLUDED>::store_hit(p, i, h);
      }
    }
  }
  if (STATS) {
    unsigned long long rays = active ? 1 : 0, nodes = st.nodes, tris = st.tris;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rays += __shfl_xor_sync(0xFFFFFFFFu, rays, o);
      nodes += __shfl_xor_sync(0xFFFFFFFFu, nodes, o);
      tris += __shfl_xor_sync(0xFFFFFFFFu, tris, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&p.stat[0], rays); atomicAdd(&p.stat[1], nodes); atomicAdd(&p.stat[2], tris);
    }
  }
}

template <int K, bool OCCLUDED>
static int launch_k(const TraceParams& p, cudaStream_t st) {
  const unsigned long long blocks = (p.n + TRACE_THREADS - 1) / TRACE_THREADS;
  if (blocks == 0) return 0;
  if (blocks > 0x7FFFFFFFull) return (int)cudaErrorInvalidValue;
  if (p.stat) trace_kernel<K, OCCLUDED, true><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(p);
  else trace_kernel<K, OCCLUDED, false><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(p);
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
