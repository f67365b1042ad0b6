// Part 1 — point -> 3D-box membership for sm_100a.
//
// Contract (bit-exact with the CPU op): SURVEY.md Appendix A.1; replaces
// mmcv.ops.points_in_boxes_{all,part,cpu} (re-exported at
// /root/reference/mmdet3d/ops/__init__.py:12-13, called from
// /root/reference/mmdet3d/core/bbox/structures/base_box3d.py:534,566).
//
// Design: the brute-force test is FP32-issue bound (14 instr x N x T), the
// output is HBM bound (16 B in, 4*W B out per point).  To sit on the HBM roofline each
// persistent CTA (one per SM) first builds, in shared memory,
//   (1) the per-box derived terms (centre z, half extents, cos/sin of -rz evaluated in
//       double with the deterministic routine of include/gga_detmath.h), and
//   (2) a BEV cull grid: for every cell a bit mask of the boxes whose conservatively
//       inflated bounding rectangle touches the cell,
// then streams its slice of points: one lane per (point, 256-box group), cell lookup, exact
// test only for the candidate bits, and 32 B of packed mask per lane written as two fully
// coalesced 512 B warp stores after a shuffle transpose.  Culling never changes the result:
// a box is a candidate wherever a point could pass the exact fp32 test.
#include <float.h>

#include "../../include/gga_detmath.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxGrid = 128;  // cells per side (7 bits in the packed range)

struct PibParams {
  const float* points;
  const float* boxes;
  void* out;
  long long items_per_frame;  // num_points * groups
  int pts_stride;
  int num_points;
  int num_boxes;
  int row_words;  // words per point row (Wp)
  int groups;     // lanes per point (row_words / WL)
  int G;          // cull grid cells per side
  int vec4;       // points are 16 B aligned float4
};

struct BoxPrep {
  float cx, cy, cz, hz;      // centre (z already shifted to the box centre), z half extent
  float cosa, sina, hx, hy;  // cos/sin of -rz, x/y half extents
};

// Per-box terms of the contract, each rounded exactly as the CPU op rounds it.
//   cz  = (float)((double)z + (double)dz / 2.0)
//   |pz - cz| > dz/2.0  (double compare)  <=>  |pz - cz| > RD_f32(dz/2)
//   lx  <  dx/2.0       (double compare)  <=>  lx <  RU_f32(dx/2),  lx > -dx/2.0 <=> lx > -RU_f32(dx/2)
// (dz/2 is exact in fp32 unless dz is subnormal; the directed roundings make the fp32
//  compares equal to the double ones in that case too.)
__device__ __forceinline__ BoxPrep prep_box(const float* __restrict__ b) {
  const float x = b[0], y = b[1], z = b[2], dx = b[3], dy = b[4], dz = b[5], rz = b[6];
  BoxPrep p;
  const double hzd = (double)dz / 2.0;
  p.cx = x;
  p.cy = y;
  p.cz = __double2float_rn(__dadd_rn((double)z, hzd));
  p.hz = __double2float_rd(hzd);
  p.hx = __double2float_ru((double)dx / 2.0);
  p.hy = __double2float_ru((double)dy / 2.0);
  double s, c;
  gga_sincos_f32(-rz, &s, &c);
  p.cosa = __double2float_rn(c);
  p.sina = __double2float_rn(s);
  return p;
}

// The exact fp32 test of the contract: separate roundings for each product and sum (the
// CPU op is built without FMA contraction), closed z slab, open x/y faces.
__device__ __forceinline__ bool inside_box(float x, float y, float z, const float4 a, const float4 r) {
  if (fabsf(__fsub_rn(z, a.z)) > a.w) return false;  // NaN z passes, like the CPU op
  const float sx = __fsub_rn(x, a.x), sy = __fsub_rn(y, a.y);
  const float lx = __fadd_rn(__fmul_rn(sx, r.x), __fmul_rn(sy, -r.y));
  const float ly = __fadd_rn(__fmul_rn(sx, r.y), __fmul_rn(sy, r.x));
  return (lx > -r.z) & (lx < r.z) & (ly > -r.w) & (ly < r.w);
}

// Conservative BEV rectangle of a box.  kind: 0 = can contain no point, 1 = finite
// rectangle, 2 = must be tested against every point (infinite extent).
// A point that passes the fp32 test has |p - c| within the rotated half extents up to a
// relative 1e-6 (rounding of the shifts, products and of cos/sin); the rectangle is
// inflated by 2^-13 of its size and every bound is rounded outwards.
__device__ __forceinline__ int box_rect(const float4 a, const float4 r, float& x0, float& x1,
                                        float& y0, float& y1) {
  const float cosa = r.x, sina = r.y, hx = r.z, hy = r.w;
  if (!(hx > 0.f) || !(hy > 0.f) || !(cosa == cosa) || !(sina == sina) || !isfinite(a.x) ||
      !isfinite(a.y))
    return 0;
  const float ac = fabsf(cosa), as = fabsf(sina);
  float ex = __fadd_ru(__fmul_ru(ac, hx), __fmul_ru(as, hy));
  float ey = __fadd_ru(__fmul_ru(as, hx), __fmul_ru(ac, hy));
  const float m = __fmul_ru(__fadd_ru(ex, ey), 1.220703125e-4f);
  ex = __fadd_ru(ex, m);
  ey = __fadd_ru(ey, m);
  x0 = __fsub_rd(a.x, ex);
  x1 = __fadd_ru(a.x, ex);
  y0 = __fsub_rd(a.y, ey);
  y1 = __fadd_ru(a.y, ey);
  if (!isfinite(x0) || !isfinite(x1) || !isfinite(y0) || !isfinite(y1)) return 2;
  return 1;
}

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct GridHdr {
  float gx0, gy0, invx, invy, fmaxx, fmaxy;
  uint32_t minx, miny, maxx, maxy;  // order-preserving encodings, reduced with atomics
  int n_rect;
  int degenerate;
};

// Monotone non-decreasing in v (one rounded subtraction, one rounded product by a
// non-negative constant): boxes and points go through the same function, so
// rect.lo <= p <= rect.hi implies cell(rect.lo) <= cell(p) <= cell(rect.hi).
__device__ __forceinline__ float fcell(float v, float g0, float inv) {
  return __fmul_rn(__fsub_rn(v, g0), inv);
}

template <int WL>
__device__ __forceinline__ void load_cand(const uint32_t* row, uint32_t (&m)[WL]) {
  if constexpr (WL == 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(row);
    const uint4 b = *reinterpret_cast<const uint4*>(row + 4);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
    m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
  } else if constexpr (WL == 4) {
    const uint4 a = *reinterpret_cast<const uint4*>(row);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
  } else if constexpr (WL == 2) {
    const uint2 a = *reinterpret_cast<const uint2*>(row);
    m[0] = a.x; m[1] = a.y;
  } else {
    m[0] = row[0];
  }
}

enum { kModeBits = 0, kModeAll = 1, kModePart = 2 };

// Dynamic shared memory layout (all 16 B aligned):
//   float4  sbox[2 * T]                  box t at [2t] = (cx, cy, cz, hz), [2t+1] = (cosa, sina, hx, hy)
//   uint32  table[(G*G + 1) * Wp]        candidate bit masks per cell (cell G*G = outside the grid)
//   uint8   summ[(G*G + 1) * groups8]    per (cell, 8-word group): which of the 8 words are non-zero
//   uint32  stage[32 warps * 256]        per-warp transpose buffer for the 32 B-per-lane stores
struct SmemLayout {
  size_t table_off, summ_off, stage_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int T, int G, int Wp, bool need_stage) {
  SmemLayout L;
  const size_t ncell = (size_t)G * G + 1;
  const size_t groups8 = (size_t)(Wp + 7) / 8;
  L.table_off = (size_t)T * 32;
  L.summ_off = L.table_off + ((ncell * Wp * 4 + 15) & ~(size_t)15);
  L.stage_off = L.summ_off + ((ncell * groups8 + 15) & ~(size_t)15);
  L.total = L.stage_off + (need_stage ? (size_t)kThreads * 32 : 0);
  return L;
}

// Cheap conservative rectangle straight from the raw box (fp32 sincosf instead of the
// double-precision contract terms): the cull grid only has to be conservative, and the
// 2^-13 inflation of box_rect dwarfs the ~1e-7 difference between sincosf and the exact
// rounded cos/sin.  This lets the grid be built while other warps are still in the long
// double-precision dependency chain of prep_box.
__device__ __forceinline__ int approx_rect(const float* __restrict__ b, float& x0, float& x1, float& y0,
                                           float& y1) {
  const float hx = __fmul_ru(b[3], 0.5f), hy = __fmul_ru(b[4], 0.5f);
  float sn, cs;
  sincosf(-b[6], &sn, &cs);
  return box_rect(make_float4(b[0], b[1], 0.f, 0.f), make_float4(cs, sn, hx, hy), x0, x1, y0, y1);
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Builds boxes + cull grid + word summaries in shared memory.
// Warp roles: the last `prep_warps` warps evaluate the exact per-box contract terms
// (double-precision chain, ~2k cycles of latency), the others build the grid concurrently.
__device__ void build_tables(const PibParams& p, const float* __restrict__ boxes, float4* sbox,
                             uint32_t* table, uint32_t* summ32, GridHdr* hdr) {
  const int T = p.num_boxes, G = p.G, Wp = p.row_words;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = G * G + 1;
  const int groups8 = (Wp + 7) >> 3;
  int prep_warps = (T + 31) >> 5;
  prep_warps = prep_warps < 1 ? 1 : (prep_warps > 16 ? 16 : prep_warps);
  const int build_warps = kThreads / 32 - prep_warps;
  const int nb = build_warps * 32;

  if (warp >= build_warps) {
    // ---- exact per-box terms (independent of the grid) ----
    for (int t = tid - nb; t < T; t += prep_warps * 32) {
      const BoxPrep q = prep_box(boxes + (long long)t * 7);
      sbox[2 * t] = make_float4(q.cx, q.cy, q.cz, q.hz);
      sbox[2 * t + 1] = make_float4(q.cosa, q.sina, q.hx, q.hy);
    }
  } else {
    // ---- cull grid ----
    if (tid == 0) {
      hdr->minx = hdr->miny = 0xffffffffu;
      hdr->maxx = hdr->maxy = 0u;
      hdr->n_rect = 0;
      hdr->degenerate = 0;
    }
    {  // zero table and summaries (contiguous, sizes padded to 16 B)
      uint4* t4 = reinterpret_cast<uint4*>(table);
      const int n4 = (int)((reinterpret_cast<unsigned char*>(summ32) - reinterpret_cast<unsigned char*>(table)) >> 4) +
                     ((ncell * groups8 + 15) >> 4);
      for (int i = tid; i < n4; i += nb) t4[i] = make_uint4(0, 0, 0, 0);
    }
    bar_sync_named(1, nb);
    for (int t = tid; t < T; t += nb) {
      float x0, x1, y0, y1;
      if (approx_rect(boxes + (long long)t * 7, x0, x1, y0, y1) == 1) {
        atomicMin(&hdr->minx, f2ord(x0));
        atomicMax(&hdr->maxx, f2ord(x1));
        atomicMin(&hdr->miny, f2ord(y0));
        atomicMax(&hdr->maxy, f2ord(y1));
        atomicAdd(&hdr->n_rect, 1);
      }
    }
    bar_sync_named(1, nb);
    if (tid == 0) {
      float gx0 = 0.f, gy0 = 0.f, invx = 0.f, invy = 0.f, fmx = -1.f, fmy = -1.f;
      if (hdr->n_rect > 0) {
        gx0 = ord2f(hdr->minx);
        gy0 = ord2f(hdr->miny);
        const float gx1 = ord2f(hdr->maxx), gy1 = ord2f(hdr->maxy);
        const float wx = gx1 - gx0, wy = gy1 - gy0;
        if (!isfinite(wx) || !isfinite(wy)) {
          hdr->degenerate = 1;  // extents overflow fp32: every box is tested against every point
        } else {
          invx = (wx > 0.f) ? (float)G / wx : 0.f;
          invy = (wy > 0.f) ? (float)G / wy : 0.f;
          if (!isfinite(invx)) invx = 0.f;
          if (!isfinite(invy)) invy = 0.f;
          fmx = fcell(gx1, gx0, invx);
          fmy = fcell(gy1, gy0, invy);
        }
      }
      hdr->gx0 = gx0; hdr->gy0 = gy0; hdr->invx = invx; hdr->invy = invy;
      hdr->fmaxx = fmx; hdr->fmaxy = fmy;
    }
    bar_sync_named(1, nb);
    // insertion: a half warp per box, lanes tile the box's cell range 4 x 4 at a time
    const float gx0 = hdr->gx0, gy0 = hdr->gy0, invx = hdr->invx, invy = hdr->invy;
    const int degenerate = hdr->degenerate;
    const float gm1 = (float)(G - 1);
    const int sub = lane >> 4, xx = lane & 3, yy = (lane >> 2) & 3;
    for (int t = warp * 2 + sub; t < T; t += build_warps * 2) {
      float x0, x1, y0, y1;
      int kind = approx_rect(boxes + (long long)t * 7, x0, x1, y0, y1);
      if (kind == 0) continue;
      if (degenerate) kind = 2;
      int cx0 = 0, cx1 = G - 1, cy0 = 0, cy1 = G - 1;
      if (kind == 1) {
        cx0 = (int)fminf(fcell(x0, gx0, invx), gm1); cx1 = (int)fminf(fcell(x1, gx0, invx), gm1);
        cy0 = (int)fminf(fcell(y0, gy0, invy), gm1); cy1 = (int)fminf(fcell(y1, gy0, invy), gm1);
      }
      const uint32_t bit = 1u << (t & 31);
      const int wi = t >> 5, g8 = t >> 8;
      const uint32_t sbit = 1u << (wi & 7);
      for (int ty = cy0 + yy; ty <= cy1; ty += 4) {
        for (int tx = cx0 + xx; tx <= cx1; tx += 4) {
          const int c = ty * G + tx;
          atomicOr(table + c * Wp + wi, bit);
          const int e = c * groups8 + g8;
          atomicOr(summ32 + (e >> 2), sbit << (8 * (e & 3)));
        }
      }
      if (kind == 2 && (lane & 15) == 0) {  // also a candidate for points outside the grid
        const int c = G * G;
        atomicOr(table + c * Wp + wi, bit);
        const int e = c * groups8 + g8;
        atomicOr(summ32 + (e >> 2), sbit << (8 * (e & 3)));
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int cell_of(float x, float y, const GridHdr& h, int G) {
  const float fx = fcell(x, h.gx0, h.invx), fy = fcell(y, h.gy0, h.invy);
  const bool in = (fx >= 0.f) & (fx <= h.fmaxx) & (fy >= 0.f) & (fy <= h.fmaxy);
  const float gm1 = (float)(G - 1);
  const int cx = (int)fminf(fx, gm1), cy = (int)fminf(fy, gm1);
  return in ? cy * G + cx : G * G;  // cell G*G: outside every finite rectangle
}

// Walks the candidate bits of one 8-word group of a cell row, one candidate per loop
// iteration (the word switch is folded into the iteration so that a warp runs
// max-over-lanes(candidates) iterations, not sum-over-words(max-over-lanes)).
// Ascending box order; `on_hit(j, b)` returns true to stop (first-hit search).
template <typename F>
__device__ __forceinline__ void walk_group(const float4* __restrict__ sbox, const uint32_t* __restrict__ row,
                                           uint32_t nz, int box_base, float x, float y, float z, F on_hit) {
  uint32_t m = 0;
  int j = 0;
  while (true) {
    if (m == 0u) {
      if (nz == 0u) break;
      j = __ffs(nz) - 1;
      nz &= nz - 1u;
      m = row[j];
    }
    const int b = __ffs(m) - 1;
    m &= m - 1u;
    const int t = box_base + j * 32 + b;
    if (inside_box(x, y, z, sbox[2 * t], sbox[2 * t + 1])) {
      if (on_hit(j, b)) break;
    }
  }
}

template <int WL, int MODE>
__global__ void __launch_bounds__(kThreads, 1) pib_kernel(const PibParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ GridHdr hdr_s;
  const int T = p.num_boxes, G = p.G, Wp = p.row_words;
  const SmemLayout lay = smem_layout(T, G, Wp, MODE == kModeBits && WL == 8);
  float4* sbox = reinterpret_cast<float4*>(smem_raw);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw + lay.table_off);
  const uint8_t* summ = smem_raw + lay.summ_off;
  const int groups8 = (Wp + 7) >> 3;

  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long items = (MODE == kModeBits) ? p.items_per_frame : (long long)p.num_points;
  const int groups = (MODE == kModeBits) ? p.groups : 1;
  const long long i0 = items * blockIdx.x / gridDim.x, i1 = items * (blockIdx.x + 1) / gridDim.x;
  // CTA-local 32-bit indexing: item li in [0, n_local) is point pt0 + (rem0 + li) / groups
  const uint32_t n_local = (uint32_t)(i1 - i0);
  const long long pt0 = i0 / groups;
  const uint32_t rem0 = (uint32_t)(i0 - pt0 * groups);
  const int gshift = (groups & (groups - 1)) == 0 ? __ffs(groups) - 1 : -1;
  const float* __restrict__ pts =
      p.points + ((long long)f * p.num_points + pt0) * p.pts_stride;  // first point of this CTA
  auto point_of = [&](uint32_t li, int& g) -> uint32_t {
    if (groups == 1) { g = 0; return li; }
    const uint32_t v = rem0 + li;
    const uint32_t q = gshift >= 0 ? (v >> gshift) : v / (uint32_t)groups;
    g = (int)(v - q * (uint32_t)groups);
    return q;
  };
  auto fetch = [&](uint32_t li, int& g, float& x, float& y, float& z) {
    if (li < n_local) {
      const uint32_t pl = point_of(li, g);
      if (p.vec4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pts) + pl);
        x = v.x; y = v.y; z = v.z;
      } else {
        const float* q = pts + (size_t)pl * p.pts_stride;
        x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2);
      }
    }
  };

  // software pipeline, distance 2: the points of the next two batches of this warp are in
  // flight while the current one is tested (issued before the table build so that the
  // first loads overlap it)
  uint32_t lb = threadIdx.x & ~31u;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f, x2 = 0.f, y2 = 0.f, z2 = 0.f;
  int g1 = 0, g2 = 0;
  fetch(lb + lane, g1, x1, y1, z1);
  fetch(lb + kThreads + lane, g2, x2, y2, z2);

  build_tables(p, p.boxes + (long long)f * T * 7, sbox, table,
               reinterpret_cast<uint32_t*>(smem_raw + lay.summ_off), &hdr_s);
  const GridHdr h = hdr_s;

  for (; lb < n_local; lb += kThreads) {
    const uint32_t li = lb + lane;
    const bool valid = li < n_local;
    const float x = x1, y = y1, z = z1;
    const int gcur = g1;
    x1 = x2; y1 = y2; z1 = z2; g1 = g2;
    fetch(li + 2 * kThreads, g2, x2, y2, z2);
    const int cell = valid ? cell_of(x, y, h, G) : G * G;

    if constexpr (MODE == kModeBits) {
      const int g = gcur;
      uint32_t w[WL];
#pragma unroll
      for (int j = 0; j < WL; ++j) w[j] = 0u;
      const uint32_t nzw = valid ? (uint32_t)summ[cell * groups8 + g] : 0u;
      walk_group(sbox, table + cell * Wp + g * WL, nzw, g * WL * 32, x, y, z, [&](int j, int b) {
#pragma unroll
        for (int k = 0; k < WL; ++k) w[k] |= (k == j) ? (1u << b) : 0u;
        return false;
      });
      uint32_t* out = reinterpret_cast<uint32_t*>(p.out) + ((long long)f * items + i0) * WL + (size_t)lb * WL;
      if constexpr (WL == 8) {
        // per-warp transpose through shared memory: lane l owns 32 B; store j writes the
        // 16 B chunk 32*j + l, so each warp store instruction covers 512 contiguous bytes
        uint4* st = reinterpret_cast<uint4*>(smem_raw + lay.stage_off) + (threadIdx.x >> 5) * 64;
        st[2 * lane] = make_uint4(w[0], w[1], w[2], w[3]);
        st[2 * lane + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        __syncwarp();
        const uint4 v0 = st[lane], v1 = st[32 + lane];
        if (lb + (lane >> 1) < n_local) __stcs(reinterpret_cast<uint4*>(out) + lane, v0);
        if (lb + 16 + (lane >> 1) < n_local) __stcs(reinterpret_cast<uint4*>(out) + 32 + lane, v1);
        __syncwarp();
      } else if constexpr (WL == 4) {
        if (valid) __stcs(reinterpret_cast<uint4*>(out) + lane, make_uint4(w[0], w[1], w[2], w[3]));
      } else if constexpr (WL == 2) {
        if (valid) __stcs(reinterpret_cast<uint2*>(out) + lane, make_uint2(w[0], w[1]));
      } else {
        if (valid) __stcs(out + lane, w[0]);
      }
    } else if constexpr (MODE == kModeAll) {
      int32_t* out = reinterpret_cast<int32_t*>(p.out) + ((long long)f * p.num_points + i0 + lb) * T;
      const int nvalid = (int)min(32u, n_local - lb);
      for (int g0 = 0; g0 < Wp; g0 += WL) {
        uint32_t w[WL];
#pragma unroll
        for (int j = 0; j < WL; ++j) w[j] = 0u;
        uint32_t nzw = 0u;
        if (valid) nzw = WL == 8 ? (uint32_t)summ[cell * groups8 + (g0 >> 3)] : (uint32_t)summ[cell];
        walk_group(sbox, table + cell * Wp + g0, nzw, g0 * 32, x, y, z, [&](int j, int b) {
#pragma unroll
          for (int k = 0; k < WL; ++k) w[k] |= (k == j) ? (1u << b) : 0u;
          return false;
        });
        // expand: for every point of the warp, lane l writes box (g0 + j) * 32 + l
#pragma unroll
        for (int j = 0; j < WL; ++j) {
          const int t = (g0 + j) * 32 + lane;
          if ((g0 + j) * 32 < T) {
            for (int q = 0; q < nvalid; ++q) {
              const uint32_t word = __shfl_sync(0xffffffffu, w[j], q);
              if (t < T) __stcs(out + (long long)q * T + t, (int32_t)((word >> lane) & 1u));
            }
          }
        }
      }
    } else {  // kModePart: first enclosing box, ascending
      if (valid) {
        int idx = -1;
        for (int g = 0; g < groups8 && idx < 0; ++g) {
          walk_group(sbox, table + cell * Wp + g * 8, (uint32_t)summ[cell * groups8 + g], g * 256, x, y, z,
                     [&](int j, int b) {
                       idx = g * 256 + j * 32 + b;
                       return true;
                     });
        }
        __stcs(reinterpret_cast<int32_t*>(p.out) + (long long)f * p.num_points + i0 + li, idx);
      }
    }
  }
}

__global__ void sincos_test_kernel(const float* __restrict__ x, long long n, float* sn, float* cs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double s, c;
    gga_sincos_f32(x[i], &s, &c);
    sn[i] = __double2float_rn(s);
    cs[i] = __double2float_rn(c);
  }
}

__global__ void box_prep_test_kernel(const float* __restrict__ boxes, int T, float* prep) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    const BoxPrep q = prep_box(boxes + (long long)t * 7);
    float* o = prep + (long long)t * 8;
    o[0] = q.cx; o[1] = q.cy; o[2] = q.cz; o[3] = q.hz;
    o[4] = q.cosa; o[5] = q.sina; o[6] = q.hx; o[7] = q.hy;
  }
}

int g_tune_grid = 0, g_tune_ctas = 0;

template <int WL, int MODE>
int launch_pib(const PibParams& p, int B, int ctas_per_frame, size_t smem, cudaStream_t st) {
  static int configured_smem[64];  // per device, grows monotonically
  int dev = 0;
  GGA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && (int)smem > configured_smem[dev]) {
    GGA_CHECK_CUDA(cudaFuncSetAttribute(pib_kernel<WL, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_smem[dev] = (int)smem;
  }
  dim3 grid(ctas_per_frame, B);
  pib_kernel<WL, MODE><<<grid, kThreads, smem, st>>>(p);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

int run_pib(int mode, const float* points, int pts_stride, const float* boxes, void* out, int B,
            int num_points, int num_boxes, void* stream) {
  GGA_REQUIRE(B >= 0 && num_points >= 0 && num_boxes >= 0, "negative size");
  GGA_REQUIRE(pts_stride >= 3, "pts_stride must be >= 3 (got %d)", pts_stride);
  if (B == 0 || num_points == 0) return GGA_OK;
  cudaStream_t st = gga_stream(stream);
  if (num_boxes == 0) {
    if (mode == kModePart) {  // every point is in no box
      GGA_REQUIRE(out, "null out pointer");
      GGA_CHECK_CUDA(cudaMemsetAsync(out, 0xff, (size_t)B * num_points * sizeof(int32_t), st));
    }
    return GGA_OK;  // bits / all have zero-width rows (the output buffer is empty)
  }
  GGA_REQUIRE(points && out, "null points/out pointer");
  GGA_REQUIRE(boxes, "null boxes pointer");
  GGA_REQUIRE(B <= 65535, "at most 65535 frames per call (got %d)", B);

  PibParams p;
  p.points = points; p.boxes = boxes; p.out = out;
  p.pts_stride = pts_stride; p.num_points = num_points; p.num_boxes = num_boxes;
  p.row_words = gga_pib_row_words(num_boxes);
  const int WL = p.row_words >= 8 ? 8 : p.row_words;
  p.groups = p.row_words / WL;
  p.items_per_frame = (long long)num_points * p.groups;
  p.vec4 = (pts_stride == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0) ? 1 : 0;

  // shared memory: boxes + cull grid + summaries (+ transpose buffer); pick the finest grid
  // that fits, capped by the automatic / tuned resolution
  const size_t max_smem = (size_t)gga_max_smem_optin() - 1024;  // static smem + slack
  const bool need_stage = (mode == kModeBits && WL == 8);
  if (smem_layout(num_boxes, 2, p.row_words, need_stage).total > max_smem) {
    gga_set_error("num_boxes=%d exceeds the shared-memory capacity of this build", num_boxes);
    return GGA_ERR_UNSUPPORTED;
  }
  int want = g_tune_grid > 0 ? g_tune_grid : 40;
  if (want < 1) want = 1;
  if (want > kMaxGrid) want = kMaxGrid;
  int G = want;
  while (G > 2 && smem_layout(num_boxes, G, p.row_words, need_stage).total > max_smem) --G;
  p.G = G;
  const size_t smem = smem_layout(num_boxes, G, p.row_words, need_stage).total;

  int ctas = g_tune_ctas > 0 ? g_tune_ctas : gga_sm_count() / B;
  if (ctas < 1) ctas = 1;
  const long long items = (mode == kModeBits) ? p.items_per_frame : (long long)num_points;
  const long long max_ctas = (items + kThreads - 1) / kThreads;
  if (ctas > max_ctas) ctas = (int)max_ctas;

#define GGA_DISPATCH(WLV)                                                                  \
  do {                                                                                     \
    if (mode == kModeBits) return launch_pib<WLV, kModeBits>(p, B, ctas, smem, st);        \
    if (mode == kModeAll) return launch_pib<WLV, kModeAll>(p, B, ctas, smem, st);          \
    return launch_pib<WLV, kModePart>(p, B, ctas, smem, st);                               \
  } while (0)
  switch (WL) {
    case 1: GGA_DISPATCH(1);
    case 2: GGA_DISPATCH(2);
    case 4: GGA_DISPATCH(4);
    default: GGA_DISPATCH(8);
  }
#undef GGA_DISPATCH
}

}  // namespace

extern "C" int gga_pib_row_words(int num_boxes) {
  if (num_boxes <= 0) return 0;
  if (num_boxes <= 32) return 1;
  if (num_boxes <= 64) return 2;
  if (num_boxes <= 128) return 4;
  return 8 * ((num_boxes + 255) / 256);
}

extern "C" int gga_pib_set_tuning(int grid_cells, int ctas_per_frame) {
  g_tune_grid = grid_cells;
  g_tune_ctas = ctas_per_frame;
  return GGA_OK;
}

extern "C" int gga_points_in_boxes_bits(const float* points, int pts_stride, const float* boxes,
                                        uint32_t* bits, int B, int num_points, int num_boxes,
                                        void* stream) {
  return run_pib(kModeBits, points, pts_stride, boxes, bits, B, num_points, num_boxes, stream);
}

extern "C" int gga_points_in_boxes_all(const float* points, int pts_stride, const float* boxes,
                                       int32_t* out, int B, int num_points, int num_boxes,
                                       void* stream) {
  return run_pib(kModeAll, points, pts_stride, boxes, out, B, num_points, num_boxes, stream);
}

extern "C" int gga_points_in_boxes_part(const float* points, int pts_stride, const float* boxes,
                                        int32_t* out, int B, int num_points, int num_boxes,
                                        void* stream) {
  return run_pib(kModePart, points, pts_stride, boxes, out, B, num_points, num_boxes, stream);
}

extern "C" int gga_points_in_boxes_all_host(const float* points, int pts_stride, const float* boxes,
                                            int32_t* out, int B, int num_points, int num_boxes) {
  GGA_REQUIRE(B >= 0 && num_points >= 0 && num_boxes >= 0, "negative size");
  if (B == 0 || num_points == 0 || num_boxes == 0) return GGA_OK;
  GGA_REQUIRE(points && boxes && out, "null pointer");
  const size_t pb = (size_t)B * num_points * pts_stride * sizeof(float);
  const size_t bb = (size_t)B * num_boxes * 7 * sizeof(float);
  const size_t ob = (size_t)B * num_points * num_boxes * sizeof(int32_t);
  float *dp = nullptr, *db = nullptr;
  int32_t* dout = nullptr;
  cudaStream_t st;
  GGA_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = GGA_OK;
  cudaError_t e = cudaMallocAsync(&dp, pb, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&db, bb, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&dout, ob, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dp, points, pb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, boxes, bb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = run_pib(kModeAll, dp, pts_stride, db, dout, B, num_points, num_boxes, st);
    if (rc == GGA_OK) e = cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, st);
  }
  if (dp) cudaFreeAsync(dp, st);
  if (db) cudaFreeAsync(db, st);
  if (dout) cudaFreeAsync(dout, st);
  const cudaError_t e2 = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (rc != GGA_OK) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) {
    gga_set_error("points_in_boxes_all_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return GGA_ERR_CUDA;
  }
  return GGA_OK;
}

extern "C" int gga_test_sincos(const float* x, int64_t n, float* sn, float* cs, void* stream) {
  if (n <= 0) return GGA_OK;
  GGA_REQUIRE(x && sn && cs, "null pointer");
  sincos_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, gga_stream(stream)>>>(x, n, sn, cs);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

extern "C" int gga_test_box_prep(const float* boxes, int num_boxes, float* prep, void* stream) {
  if (num_boxes <= 0) return GGA_OK;
  GGA_REQUIRE(boxes && prep, "null pointer");
  box_prep_test_kernel<<<(num_boxes + 127) / 128, 128, 0, gga_stream(stream)>>>(boxes, num_boxes, prep);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
