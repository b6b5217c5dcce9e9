// rt_core_h2.cuh -- PARKED EXPERIMENT (test tool, not part of the product): a packed-half version of the BVH8 slab test.
//
// Idea: evaluate the child planes in HALF precision *relative to the ray's entry time t0 into the node's grid box*, scaled
// by a power of two so that the longest axis crossing maps to < 2^15, two children per HFMA2 / HMNMX2.  Relative to t0
// every plane distance of the binding axes is at most one node crossing, so a half resolves 1/8 of a grid cell; the three
// roundings involved are covered by moving every entry plane down and every exit plane up by HALF A CELL.  The test is a
// superset of the fp32 test: the CPU emulation (this file compiled with g++, _Float16 arithmetic bit-identical to the
// device ops: round-to-nearest conversions, fused HFMA2, exact HSUB2 of 1024) reports bit-identical hits and ~3 % more
// node visits (profiles/r1_h16_model.txt).
// Outcome: compiled into trace.cu for sm_100a (nvcc 12.9) the kernel grew from 800 to 872 SASS instructions -- the 48
// I2F + 48 FFMA + 34 FMNMX became 24 PRMT + 24 HADD2 + 24 HFMA2 + 16 HMNMX2, but the t0 / scale / pad / pack prologue
// (+45) and the per-child hit-mask assembly, which the packing does not touch (+16 for turning HSET2 masks back into
// predicates), ate the saving.  Not worth a GPU run in this form; kept for the next attempt (vectorised mask assembly).
#pragma once
#include "../../embree_b200/csrc/rt_core.cuh"

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define RT_H2 __device__ __forceinline__
namespace rtk {
typedef __half2 H2;
RT_H2 H2 h2_bcast(float a) { return __float2half2_rn(a); }
RT_H2 H2 h2_fma(H2 a, H2 b, H2 c) { return __hfma2(a, b, c); }
RT_H2 H2 h2_sub(H2 a, H2 b) { return __hsub2(a, b); }
RT_H2 H2 h2_max(H2 a, H2 b) { return __hmax2(a, b); }
RT_H2 H2 h2_min(H2 a, H2 b) { return __hmin2(a, b); }
RT_H2 uint32_t h2_le_mask(H2 a, H2 b) { return __hle2_mask(a, b); }          // 0xFFFF per half where a <= b
// bytes j and j+1 of x as the half pair (1024 + x.byte[j], 1024 + x.byte[j+1]): 0x6400 | q is exactly 1024 + q
RT_H2 H2 h2_bytes_1024(uint32_t x, int j) {
  const uint32_t v = __byte_perm(x, 0x64646464u, 0x4040u | (uint32_t)j | ((uint32_t)(j + 1) << 8));
  return *reinterpret_cast<const H2*>(&v);
}
}  // namespace rtk
#else
#define RT_H2 inline
namespace rtk {
struct H2 { _Float16 x, y; };
RT_H2 H2 h2_bcast(float a) { return H2{(_Float16)a, (_Float16)a}; }
RT_H2 _Float16 h1_fma(_Float16 a, _Float16 b, _Float16 c) { return (_Float16)((double)a * (double)b + (double)c); }
RT_H2 H2 h2_fma(H2 a, H2 b, H2 c) { return H2{h1_fma(a.x, b.x, c.x), h1_fma(a.y, b.y, c.y)}; }
RT_H2 H2 h2_sub(H2 a, H2 b) { return H2{(_Float16)((float)a.x - (float)b.x), (_Float16)((float)a.y - (float)b.y)}; }
RT_H2 H2 h2_max(H2 a, H2 b) { return H2{a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y}; }
RT_H2 H2 h2_min(H2 a, H2 b) { return H2{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y}; }
RT_H2 uint32_t h2_le_mask(H2 a, H2 b) { return (a.x <= b.x ? 0xFFFFu : 0u) | (a.y <= b.y ? 0xFFFF0000u : 0u); }
RT_H2 H2 h2_bytes_1024(uint32_t x, int j) {
  return H2{(_Float16)(1024.0f + (float)((x >> (8 * j)) & 0xFFu)), (_Float16)(1024.0f + (float)((x >> (8 * j + 8)) & 0xFFu))};
}
}  // namespace rtk
#endif

namespace rtk {

// Same contract as node_hitmask() (rt_core.cuh): returns [31:24] internal children by traversal priority, [23:0] one
// bit per triangle of the node's leaf slots, for a superset of the children the fp32 test accepts.
RT_H2 uint32_t node_hitmask_h2(const u32x4& n0, const u32x4& n1, const u32x4& n2, const u32x4& n3, const u32x4& n4, float ox,
                               float oy, float oz, float idx, float idy, float idz, bool negx, bool negy, bool negz, float tnear,
                               float tfar, uint32_t oct_inv4) {
  const uint32_t e = n0.w;
  const float sx = u2f((e & 0xFFu) << 23) * idx, sy = u2f(((e >> 8) & 0xFFu) << 23) * idy, sz = u2f(((e >> 16) & 0xFFu) << 23) * idz;
  const float bx = (u2f(n0.x) - ox) * idx, by = (u2f(n0.y) - oy) * idy, bz = (u2f(n0.z) - oz) * idz;
  // entry time into the node's grid box (all 255 cells per axis), never before tnear
  const float t0 = fmaxf(fmaxf(fminf(bx, fma_rn(255.0f, sx, bx)), fminf(by, fma_rn(255.0f, sy, by))),
                         fmaxf(fminf(bz, fma_rn(255.0f, sz, bz)), tnear));
  // power-of-two scale: the longest axis crossing (255 cells) lands in [2^13.99, 2^15)
  const float smax = fmaxf(fmaxf(fabsf(sx), fabsf(sy)), fabsf(sz));
  uint32_t eb = (f2u(smax) >> 23) & 0xFFu;
  eb = eb < 7u ? 7u : eb;
  const float k = u2f((260u - eb) << 23);
  const float slack = mul_rn(mul_rn(fabsf(t0), 4.8e-7f), k);   // fp32 rounding of b - t0 (2 ulp of t0), in scaled units
  const float kLim = 60000.0f;
  const float skx = mul_rn(sx, k), sky = mul_rn(sy, k), skz = mul_rn(sz, k);
  const float bkx = mul_rn(sub_rn(bx, t0), k), bky = mul_rn(sub_rn(by, t0), k), bkz = mul_rn(sub_rn(bz, t0), k);
  const float px = fma_rn(0.5f, fabsf(skx), slack), py = fma_rn(0.5f, fabsf(sky), slack), pz = fma_rn(0.5f, fabsf(skz), slack);
  const H2 shx = h2_bcast(skx), shy = h2_bcast(sky), shz = h2_bcast(skz);
  const H2 bnx = h2_bcast(fminf(fmaxf(sub_rn(bkx, px), -kLim), kLim)), bfx = h2_bcast(fminf(fmaxf(add_rn(bkx, px), -kLim), kLim));
  const H2 bny = h2_bcast(fminf(fmaxf(sub_rn(bky, py), -kLim), kLim)), bfy = h2_bcast(fminf(fmaxf(add_rn(bky, py), -kLim), kLim));
  const H2 bnz = h2_bcast(fminf(fmaxf(sub_rn(bkz, pz), -kLim), kLim)), bfz = h2_bcast(fminf(fmaxf(add_rn(bkz, pz), -kLim), kLim));
  // ray interval relative to t0, widened by a half ulp-and-a-bit in each direction
  const float tnk = fminf(fmaxf(mul_rn(sub_rn(tnear, t0), k), -kLim), kLim), tfk = fminf(fmaxf(mul_rn(sub_rn(tfar, t0), k), -kLim), kLim);
  const H2 tn2 = h2_bcast(sub_rn(tnk, fma_rn(fabsf(tnk), 0.001f, 1e-6f))), tf2 = h2_bcast(add_rn(tfk, fma_rn(fabsf(tfk), 0.001f, 1e-6f)));
  const H2 k1024 = h2_bcast(1024.0f);
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t meta4 = half ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
    const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
    const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
    const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
    const uint32_t nz = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#pragma unroll
    for (int j = 0; j < 4; j += 2) {   // children j and j+1 of this half in the two halves of every H2
      const H2 tnx = h2_fma(h2_sub(h2_bytes_1024(nx, j), k1024), shx, bnx);
      const H2 tny = h2_fma(h2_sub(h2_bytes_1024(ny, j), k1024), shy, bny);
      const H2 tnz = h2_fma(h2_sub(h2_bytes_1024(nz, j), k1024), shz, bnz);
      const H2 tfx = h2_fma(h2_sub(h2_bytes_1024(fx, j), k1024), shx, bfx);
      const H2 tfy = h2_fma(h2_sub(h2_bytes_1024(fy, j), k1024), shy, bfy);
      const H2 tfz = h2_fma(h2_sub(h2_bytes_1024(fz, j), k1024), shz, bfz);
      const H2 tmin = h2_max(h2_max(tnx, tny), h2_max(tnz, tn2));
      const H2 tmax = h2_min(h2_min(tfx, tfy), h2_min(tfz, tf2));
      const uint32_t le = h2_le_mask(tmin, tmax);
      if (le & 0x0000FFFFu) hitmask |= ((child_bits4 >> (8 * j)) & 0xFFu) << ((bit_index4 >> (8 * j)) & 0xFFu);
      if (le & 0xFFFF0000u) hitmask |= ((child_bits4 >> (8 * j + 8)) & 0xFFu) << ((bit_index4 >> (8 * j + 8)) & 0xFFu);
    }
  }
  return hitmask;
}

}  // namespace rtk
