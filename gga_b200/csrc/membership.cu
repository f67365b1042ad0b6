// Part 1 — point -> 3D-box membership for sm_100a.
//
// Contract (bit-exact with the CPU op): SURVEY.md Appendix A.1; replaces
// mmcv.ops.points_in_boxes_{all,part,cpu} (re-exported at
// /root/reference/mmdet3d/ops/__init__.py:12-13, called from
// /root/reference/mmdet3d/core/bbox/structures/base_box3d.py:534,566).
//
// Design.  The brute-force test is FP32-issue bound (14 instr x N x T); the
// useful traffic is 16 B in and 4*W B out per point.  To sit on the HBM roofline the work is
// split in two launches chained with programmatic dependent launch (PDL):
//
//  1. pib_prep_kernel — once per frame, a few CTAs per frame: the per-box contract terms
//     (centre z, half extents, cos/sin of -rz from the deterministic double routine of
//     include/gga_detmath.h) and a BEV "frame index": a G x G grid whose 32-bit cell word is
//     empty / one candidate box id / (count, offset) into a packed id list.  Boxes are
//     rasterised conservatively (AABB of the inflated rotated footprint refined by a
//     separating-axis test per cell), so a box is listed wherever a point could pass the
//     exact fp32 test; G is chosen per frame so that the lists fit their pool.
//  2. pib_stream_kernel — a persistent grid of light CTAs (256 threads, 6 per SM) streams the
//     points: one LDG.128 per point, one L1-cached cell-word lookup, the exact test only for
//     the listed candidates, and the packed mask rows staged per warp in shared memory so
//     that every warp store is a fully coalesced 512 B STG.128.  Warps take 4-batch tiles
//     from per-range counters (ranges follow blockIdx % R so that CTAs sharing an SM share a
//     frame's index in L1) and steal from neighbouring ranges at the end.
//
// Culling never changes the result; every candidate is decided by inside_box() below.
#include <float.h>

#include "../../include/gga_detmath.h"
#include "common.cuh"

namespace {

#ifndef GGA_STREAM_THREADS
#define GGA_STREAM_THREADS 256
#endif
#ifndef GGA_OCC
#define GGA_OCC 5
#endif
constexpr int kStreamThreads = GGA_STREAM_THREADS;
constexpr int kWarps = kStreamThreads / 32;
constexpr int kOcc = GGA_OCC;  // stream CTAs per SM (5 x 256 threads: 48 registers per thread)
constexpr int kPrepThreads = 512;
constexpr int kMaxRanges = 1024;
constexpr int kMaxSliceCells = 4096;
constexpr int kMaxG = 256;
// Cell word of the frame index (32 bits):
//   0                      empty
//   01 | 0 | id0           one candidate            (ids are 15 bits)
//   10 | id1 | id0         two candidates
//   11 | cnt:10 | off:20   list of cnt ids at ids[off]; cnt == 1023: real count is ids[off], list at off + 1
//   0xffffffff             every box is a candidate (id pool overflow fallback)
constexpr uint32_t kKindShift = 30;
constexpr uint32_t kIdMask = 0x7fffu;
constexpr uint32_t kListCntShift = 20;
constexpr uint32_t kListOffMask = (1u << kListCntShift) - 1u;
constexpr uint32_t kCntLong = 1023u;
constexpr uint32_t kCellAll = 0xffffffffu;
constexpr uint32_t kMaxCap = kListOffMask - 1u;
constexpr int kInline = 4;  // candidates a cell can hold in shared memory during the single raster pass
constexpr int kMaxBoxes = 32767;

enum { kModeBits = 0, kModeAll = 1, kModePart = 2 };

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
__device__ __forceinline__ int box_rect(float cx, float cy, float cosa, float sina, float hx, float hy,
                                        float& x0, float& x1, float& y0, float& y1) {
  if (!(hx > 0.f) || !(hy > 0.f) || !(cosa == cosa) || !(sina == sina) || !isfinite(cx) || !isfinite(cy))
    return 0;
  const float ac = fabsf(cosa), as = fabsf(sina);
  float ex = __fadd_ru(__fmul_ru(ac, hx), __fmul_ru(as, hy));
  float ey = __fadd_ru(__fmul_ru(as, hx), __fmul_ru(ac, hy));
  const float m = __fmul_ru(__fadd_ru(ex, ey), 1.220703125e-4f);
  ex = __fadd_ru(ex, m);
  ey = __fadd_ru(ey, m);
  x0 = __fsub_rd(cx, ex);
  x1 = __fadd_ru(cx, ex);
  y0 = __fsub_rd(cy, ey);
  y1 = __fadd_ru(cy, ey);
  if (!isfinite(x0) || !isfinite(x1) || !isfinite(y0) || !isfinite(y1)) return 2;
  return 1;
}

// What the rasteriser needs of a box.  fp32 sincosf instead of the double-precision contract
// terms: the index only has to be conservative, and the 2^-13 inflation dwarfs the ~1e-7
// difference between sincosf and the exactly rounded cos/sin.
struct RasterBox {
  int kind;
  float cx, cy, cs, sn, hx, hy;
  float x0, x1, y0, y1;
};

__device__ __forceinline__ RasterBox raster_box(const float* __restrict__ b) {
  RasterBox r;
  r.cx = b[0];
  r.cy = b[1];
  r.hx = __fmul_ru(b[3], 0.5f);
  r.hy = __fmul_ru(b[4], 0.5f);
  sincosf(-b[6], &r.sn, &r.cs);
  r.x0 = r.x1 = r.y0 = r.y1 = 0.f;
  r.kind = box_rect(r.cx, r.cy, r.cs, r.sn, r.hx, r.hy, r.x0, r.x1, r.y0, r.y1);
  return r;
}

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Per-frame header of the index (48 B, written by slice 0 of the prep kernel).
// The grid is (G + 2) x (G + 2): the inner G x G cells tile the bounding rectangle of all
// finite footprints, the one-cell border catches everything outside it (and NaN), so points
// and footprints go through ONE clamped, monotone cell function and no case is special.
struct __align__(16) FrameHdr {
  float gx0, gy0, invx, invy;  // padded cell = clamp(floor((v - g0) * inv + 1), 0, G + 1)
  float cwx, cwy;              // inner cell size in metres (inf when the extent is degenerate)
  float slopx, slopy;          // absolute slack of the cell geometry (rounding of the cell function)
  int G, n_rect, n_inf, pad;
};

// Monotone non-decreasing in v: one rounded subtraction, one rounded product by a
// non-negative constant, one rounded addition, clamps, truncation.  Boxes and points go
// through the same function, so rect.lo <= p <= rect.hi implies cell(rect.lo) <= cell(p) <=
// cell(rect.hi).  NaN maps to the last cell (fminf returns the non-NaN operand).
__device__ __forceinline__ int pcell(float v, float g0, float inv, float gp1) {
  const float f = __fadd_rn(__fmul_rn(__fsub_rn(v, g0), inv), 1.0f);
  return (int)fmaxf(fminf(f, gp1), 0.f);
}

// Workspace layout (device memory owned by the caller, see gga_pib_workspace_bytes):
//   FrameHdr hdr[F]
//   float4   prep[F][2 T]       box t: [2t] = (cx, cy, cz, hz), [2t+1] = (cosa, sina, hx, hy)
//   uint32   grid[F][gstride]   cell words, row-major (G + 2) x (G + 2) of the frame's own G
//   uint16   ids[F][cap]        every slice of the prep grid owns cap / S entries
struct WsLayout {
  size_t hdr, prep, grid, ids, total;
  size_t gstride;  // words per frame
  uint32_t cap;    // ids per frame
  int Gmax;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline WsLayout ws_layout(int F, int T, int Gmax) {
  WsLayout L;
  L.Gmax = Gmax;
  L.gstride = align_up((size_t)(Gmax + 2) * (Gmax + 2), 4);
  size_t cap = (size_t)8 * Gmax * Gmax;
  if (cap < (size_t)16 * T + 64) cap = (size_t)16 * T + 64;
  if (cap > kMaxCap) cap = kMaxCap;
  cap = cap & ~(size_t)7;
  L.cap = (uint32_t)cap;
  size_t o = 0;
  L.hdr = o; o = align_up(o + (size_t)F * sizeof(FrameHdr), 256);
  L.prep = o; o = align_up(o + (size_t)F * T * 32, 256);
  L.grid = o; o = align_up(o + (size_t)F * L.gstride * 4, 256);
  L.ids = o; o = align_up(o + (size_t)F * cap * 2, 256);
  L.total = o;
  return L;
}

int g_tune_grid = 0, g_tune_occ = 0, g_tune_phase = 0, g_tune_nogsm = 0, g_tune_nofast = 0, g_tune_gsm = 0;
unsigned long long* g_trace = nullptr;
unsigned long long* g_trace_prep = nullptr;

inline int pick_gmax(int N, int T) {
  if (g_tune_grid > 0) return g_tune_grid > kMaxG ? kMaxG : g_tune_grid;
  double cells = 64.0 * (double)T;
  if (cells > 4.0 * (double)N) cells = 4.0 * (double)N;
  int G = (int)(sqrt(cells) + 0.5);
  if (G < 4) G = 4;
  if (G > 192) G = 192;
  return G;
}

// ------------------------------------------------------------------------------------------
// prep kernel
// ------------------------------------------------------------------------------------------
struct PrepParams {
  const float* boxes;  // [F, T, 7]
  unsigned char* ws;
  WsLayout L;
  int T;
  int ncache;  // boxes whose raster terms are cached in shared memory
  int Gcap;             // largest G whose slices fit the shared-memory cell arrays
  int S;                // index slices per frame (grid.x = S + CTAs for the exact contract terms)
  int max_slice_cells;  // cells of a slice at G = Gcap
  unsigned long long* trace;  // profiling hook: 16 globaltimer stamps per CTA (NULL = off)
};

__device__ __forceinline__ void cell_range(const RasterBox& rb, const FrameHdr& h, int& cx0, int& cx1, int& cy0,
                                           int& cy1) {
  const int G = h.G;
  cx0 = 0; cy0 = 0; cx1 = G + 1; cy1 = G + 1;
  if (rb.kind == 1) {
    const float gp1 = (float)(G + 1);
    cx0 = pcell(rb.x0, h.gx0, h.invx, gp1);
    cx1 = pcell(rb.x1, h.gx0, h.invx, gp1);
    cy0 = pcell(rb.y0, h.gy0, h.invy, gp1);
    cy1 = pcell(rb.y1, h.gy0, h.invy, gp1);
  }
}

constexpr int kBoxWords = 11;

// Per-(box, grid) constants of the separating-axis test, hoisted out of the cell loop.
struct SatBox {
  float cx, cy, cs, sn, bx, by;
};
__device__ __forceinline__ SatBox sat_box(const RasterBox& rb, const FrameHdr& h) {
  SatBox s;
  s.cx = rb.cx; s.cy = rb.cy; s.cs = rb.cs; s.sn = rb.sn;
  const float Hx = 0.5f * h.cwx + h.slopx, Hy = 0.5f * h.cwy + h.slopy;
  const float ac = fabsf(rb.cs), as = fabsf(rb.sn);
  const float infl = (rb.hx + rb.hy + Hx + Hy) * 1.220703125e-4f;
  s.bx = rb.hx + ac * Hx + as * Hy + infl;
  s.by = rb.hy + as * Hx + ac * Hy + infl;
  return s;
}
// Conservative "could a point of the INNER cell (tx, ty) (padded coordinates 1..G) pass the
// exact test of this box": separating-axis test in the box frame.
__device__ __forceinline__ bool sat_overlap(const SatBox& s, const FrameHdr& h, int tx, int ty) {
  const float dx = h.gx0 + ((float)tx - 0.5f) * h.cwx - s.cx, dy = h.gy0 + ((float)ty - 0.5f) * h.cwy - s.cy;
  const float lx = dx * s.cs - dy * s.sn, ly = dx * s.sn + dy * s.cs;
  return !(fabsf(lx) > s.bx) && !(fabsf(ly) > s.by);  // NaN / inf anywhere -> keep the box
}

__global__ void __launch_bounds__(kPrepThreads) pib_prep_kernel(const PrepParams p) {
  // let the dependent stream kernel start its prologue right away
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int tk = 0;
  auto stamp = [&]() {
    if (p.trace && threadIdx.x == 0 && tk < 16) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + tk] = t;
    }
    ++tk;
  };
  stamp();
  extern __shared__ __align__(16) unsigned char dsm[];
  constexpr int kCells = kMaxSliceCells + 4;
  uint32_t* cf = reinterpret_cast<uint32_t*>(dsm);   // candidates per cell (raster pass), then n | fill cursor << 16
  uint32_t* word = cf + kCells;                      // inline ids 0,1
  uint32_t* word2 = word + kCells;                   // inline ids 2,3 -> list offset of cells with > kInline boxes
  uint32_t* mine = word2 + kCells;                   // [T] boxes touching this slice: id | cx0 << 16, then cx1 | cy0 << 8 ...
  float* cbox = reinterpret_cast<float*>(dsm + (size_t)3 * kCells * 4 + align_up((size_t)p.T * 12, 16));
  constexpr int kNW = kPrepThreads / 32;
  __shared__ uint32_t s_minx, s_miny, s_maxx, s_maxy;
  __shared__ int s_nrect, s_ninf, s_nmine, s_big;
  __shared__ float s_sw[kNW], s_sh[kNW], s_swh[kNW];
  __shared__ uint32_t s_cursor;

  const int T = p.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.y, slice = blockIdx.x, S = p.S;
  const float* __restrict__ boxes = p.boxes + (size_t)f * T * 7;
  unsigned char* ws = p.ws;
  float4* prep = reinterpret_cast<float4*>(ws + p.L.prep) + (size_t)f * 2 * T;
  if (slice >= S) {
    // CTAs beyond the S index slices: the exact (double precision) contract terms, one box per
    // thread.  A ~3 us dependent chain (measured) that nothing in this kernel waits for.
    const int t = (slice - S) * kPrepThreads + tid;
    if (t < T) {
      const BoxPrep q = prep_box(boxes + (size_t)t * 7);
      prep[2 * t] = make_float4(q.cx, q.cy, q.cz, q.hz);
      prep[2 * t + 1] = make_float4(q.cosa, q.sina, q.hx, q.hy);
    }
    return;
  }
  uint32_t* grid = reinterpret_cast<uint32_t*>(ws + p.L.grid) + (size_t)f * p.L.gstride;
  // this slice's own share of the frame's id pool (static split: no cursor, no zero-init contract)
  const uint32_t pool = (p.L.cap / (uint32_t)S) & ~7u;
  const uint32_t pool0 = (uint32_t)slice * pool;
  uint16_t* ids = reinterpret_cast<uint16_t*>(ws + p.L.ids) + (size_t)f * p.L.cap;
  const int ncache = p.ncache;

  auto get_box = [&](int t) -> RasterBox {
    if (t < ncache) {
      const float* c = cbox + t * kBoxWords;
      RasterBox r;
      r.x0 = c[0]; r.x1 = c[1]; r.y0 = c[2]; r.y1 = c[3];
      r.cx = c[4]; r.cy = c[5]; r.cs = c[6]; r.sn = c[7]; r.hx = c[8]; r.hy = c[9];
      r.kind = __float_as_int(c[10]);
      return r;
    }
    return raster_box(boxes + (size_t)t * 7);
  };

  // request this thread's first box before anything else: at the training shape the boxes come
  // from DRAM behind the previous step's mask write-back (~2 us measured), the longest wait here
  float b0[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) b0[j] = tid < T ? __ldg(boxes + (size_t)tid * 7 + j) : 0.f;
  if (tid == 0) {
    s_minx = s_miny = 0xffffffffu;
    s_maxx = s_maxy = 0u;
    s_nrect = s_ninf = s_nmine = s_big = 0;
    s_cursor = 0u;
  }
  // cell arrays of the largest slice this launch can produce
  for (int i = tid; i < p.max_slice_cells; i += kPrepThreads) { cf[i] = 0u; word[i] = 0u; word2[i] = 0u; }
  __syncthreads();
  stamp();

  // ---- pass A: footprints, their extent, and the sums that predict the list length ---------
  {
    float sw = 0.f, sh = 0.f, swh = 0.f;
    uint32_t mnx = 0xffffffffu, mny = 0xffffffffu, mxx = 0u, mxy = 0u;
    int nrect = 0, ninf = 0;
    for (int t = tid; t < T; t += kPrepThreads) {
      const RasterBox rb = t == tid ? raster_box(b0) : raster_box(boxes + (size_t)t * 7);
      if (t < ncache) {
        float* c = cbox + t * kBoxWords;
        c[0] = rb.x0; c[1] = rb.x1; c[2] = rb.y0; c[3] = rb.y1;
        c[4] = rb.cx; c[5] = rb.cy; c[6] = rb.cs; c[7] = rb.sn; c[8] = rb.hx; c[9] = rb.hy;
        c[10] = __int_as_float(rb.kind);
      }
      if (rb.kind == 1) {
        mnx = min(mnx, f2ord(rb.x0)); mxx = max(mxx, f2ord(rb.x1));
        mny = min(mny, f2ord(rb.y0)); mxy = max(mxy, f2ord(rb.y1));
        ++nrect;
        const float w = rb.x1 - rb.x0, hgt = rb.y1 - rb.y0;
        sw += w; sh += hgt; swh += w * hgt;
      } else if (rb.kind == 2) {
        ++ninf;
      }
    }
    // one shared-memory atomic per warp and quantity (redux.sync), not one per box
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    nrect = __reduce_add_sync(0xffffffffu, nrect); ninf = __reduce_add_sync(0xffffffffu, ninf);
    // fixed reduction tree: every slice of the frame computes bit-identical sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      sh += __shfl_xor_sync(0xffffffffu, sh, o);
      swh += __shfl_xor_sync(0xffffffffu, swh, o);
    }
    if (lane == 0) {
      atomicMin(&s_minx, mnx); atomicMax(&s_maxx, mxx);
      atomicMin(&s_miny, mny); atomicMax(&s_maxy, mxy);
      if (nrect) atomicAdd(&s_nrect, nrect);
      if (ninf) atomicAdd(&s_ninf, ninf);
      s_sw[warp] = sw; s_sh[warp] = sh; s_swh[warp] = swh;
    }
  }
  __syncthreads();
  stamp();
  // ---- header: every thread computes it redundantly from the block-wide results (same code,
  //      same inputs, same bits) — cheaper than one thread computing and a barrier publishing ----
  FrameHdr h;
  {
    float sw = lane < kNW ? s_sw[lane] : 0.f, sh = lane < kNW ? s_sh[lane] : 0.f, swh = lane < kNW ? s_swh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      sh += __shfl_xor_sync(0xffffffffu, sh, o);
      swh += __shfl_xor_sync(0xffffffffu, swh, o);
    }
    const int nrect = s_nrect, ninf = s_ninf;
    h.n_rect = nrect; h.n_inf = ninf; h.pad = 0;
    float gx0 = 0.f, gy0 = 0.f, gx1 = 0.f, gy1 = 0.f, wx = 0.f, wy = 0.f;
    bool degenerate = true;  // no finite footprint, or extents overflow fp32: one inner cell
    if (nrect > 0) {
      gx0 = ord2f(s_minx); gy0 = ord2f(s_miny);
      gx1 = ord2f(s_maxx); gy1 = ord2f(s_maxy);
      wx = gx1 - gx0;
      wy = gy1 - gy0;
      degenerate = !isfinite(wx) || !isfinite(wy);
    }
    int G = 1;
    float invx = 0.f, invy = 0.f, cwx = INFINITY, cwy = INFINITY;
    if (!degenerate) {
      // entries(G) ~ sum over boxes of (w G / Lx + 2)(h G / Ly + 2), + (G + 2)^2 per unbounded box
      const float rx = wx > 0.f ? __frcp_rn(wx) : 0.f, ry = wy > 0.f ? __frcp_rn(wy) : 0.f;
      const float ax = wx > 0.f ? sw * rx : (float)nrect, ay = wy > 0.f ? sh * ry : (float)nrect;
      const float axy = (wx > 0.f && wy > 0.f) ? swh * rx * ry : (float)nrect;
      const float budget = 0.5f * (float)p.L.cap;  // half: the pool is split statically between the slices
      G = p.Gcap;                                  // largest G whose slices fit the shared-memory cell arrays
      while (G > 1) {
        const float g = (float)G;
        const float e = axy * g * g + 2.f * (ax + ay) * g + 4.f * (float)nrect + (float)ninf * (g + 2.f) * (g + 2.f);
        if (e <= budget) break;  // (NaN e: keep shrinking)
        G = G * 7 / 8;
        if (G < 1) G = 1;
      }
      // any positive scale works (points and footprints share it); the reciprocal of the
      // estimate above is reused, its 1-ulp error is inside the cell slack of the SAT test
      invx = (float)G * rx;
      invy = (float)G * ry;
      if (!isfinite(invx)) invx = 0.f;
      if (!isfinite(invy)) invy = 0.f;
      const float rg = __frcp_rn((float)G);
      if (invx > 0.f) cwx = wx * rg;
      if (invy > 0.f) cwy = wy * rg;
    } else {
      gx0 = gy0 = gx1 = gy1 = 0.f;
    }
    h.G = G;
    h.gx0 = gx0; h.gy0 = gy0;
    h.invx = invx; h.invy = invy; h.cwx = cwx; h.cwy = cwy;
    h.slopx = (fabsf(gx0) + fabsf(gx1)) * 3.814697265625e-6f;  // 2^-18
    h.slopy = (fabsf(gy0) + fabsf(gy1)) * 3.814697265625e-6f;
  }
  const int G = h.G, Gp = G + 2;
  const int rps = (Gp + S - 1) / S;
  const int row0 = min(Gp, slice * rps), row1 = min(Gp, row0 + rps);
  const int ncell = (row1 - row0) * Gp;  // cells of this slice (<= p.max_slice_cells)
  if (slice == 0 && tid == 0) reinterpret_cast<FrameHdr*>(ws + p.L.hdr)[f] = h;

  // ---- pass C: boxes touching this slice, with their cell ranges --------------------------
  for (int t = tid; t < T; t += kPrepThreads) {
    const RasterBox rb = get_box(t);
    if (rb.kind == 0) continue;
    int cx0, cx1, cy0, cy1;
    cell_range(rb, h, cx0, cx1, cy0, cy1);
    if (cy1 >= row0 && cy0 < row1) {
      const int k = atomicAdd(&s_nmine, 1);
      mine[3 * k] = (uint32_t)t;
      mine[3 * k + 1] = (uint32_t)cx0 | ((uint32_t)cx1 << 16);
      mine[3 * k + 2] = (uint32_t)max(cy0, row0) | ((uint32_t)min(cy1, row1 - 1) << 16);
    }
  }
  __syncthreads();
  stamp();
  const int nmine = s_nmine;

  // Visits every (box, local cell) incidence of this slice; a half warp per box, its 16
  // lanes tile the box's cell range 4 x 4 at a time.  Border cells take every box whose
  // range reaches them; inner cells are refined by the separating-axis test.
  auto raster = [&](auto visit) {
    const int hw = tid >> 4, l16 = tid & 15, xx = l16 & 3, yy = l16 >> 2;
    for (int i = hw; i < nmine; i += kPrepThreads / 16) {
      const int t = (int)mine[3 * i];
      const uint32_t rx = mine[3 * i + 1], ry = mine[3 * i + 2];
      const int cx0 = (int)(rx & 0xffffu), cx1 = (int)(rx >> 16), y0 = (int)(ry & 0xffffu), y1 = (int)(ry >> 16);
      const RasterBox rb = get_box(t);
      const SatBox sb = sat_box(rb, h);
      for (int ty = y0 + yy; ty <= y1; ty += 4)
        for (int tx = cx0 + xx; tx <= cx1; tx += 4) {
          const bool border = (tx == 0) | (tx == G + 1) | (ty == 0) | (ty == G + 1);
          if (border || rb.kind == 2 || sat_overlap(sb, h, tx, ty)) visit(t, (ty - row0) * Gp + tx);
        }
    }
  };

  // ---- the raster pass: count, and keep the first kInline candidates of every cell ---------
  raster([&](int t, int c) {
    const uint32_t k = atomicAdd(&cf[c], 1u);
    if (k < 2u) atomicOr(&word[c], (uint32_t)t << (15u * k));
    else if (k < (uint32_t)kInline) atomicOr(&word2[c], (uint32_t)t << (15u * (k - 2u)));
    else if (k == (uint32_t)kInline) s_big = 1;
  });
  __syncthreads();
  stamp();

  // ---- emit: final cell words straight to the grid; list space from a slice-local cursor ---
  const bool any_big = s_big != 0;
  for (int c = tid; c < ncell; c += kPrepThreads) {
    const uint32_t n = cf[c];
    const uint32_t w01 = word[c], w23 = word2[c];
    uint32_t w = 0u;
    if (n == 1u) {
      w = (1u << kKindShift) | (w01 & kIdMask);
    } else if (n == 2u) {
      w = (2u << kKindShift) | (w01 & 0x3fffffffu);
    } else if (n >= 3u) {
      const bool lng = n >= kCntLong;
      const uint32_t need = n + (lng ? 1u : 0u);
      const uint32_t off = atomicAdd(&s_cursor, need);
      if (off + need > pool) {
        w = kCellAll;  // this slice's share of the pool is exhausted: test every box (slow, exact)
        cf[c] = 0u;
      } else {
        const uint32_t o = pool0 + off;
        w = (3u << kKindShift) | ((lng ? kCntLong : n) << kListCntShift) | o;
        if (n <= (uint32_t)kInline) {
          ids[o] = (uint16_t)(w01 & kIdMask);
          ids[o + 1] = (uint16_t)((w01 >> 15) & kIdMask);
          ids[o + 2] = (uint16_t)(w23 & kIdMask);
          if (n == 4u) ids[o + 3] = (uint16_t)((w23 >> 15) & kIdMask);
        } else {
          if (lng) ids[o] = (uint16_t)n;
          word2[c] = o + (lng ? 1u : 0u);  // where the fill pass writes this cell's list
        }
      }
    }
    grid[row0 * Gp + c] = w;
  }
  // ---- fill pass, only when some cell of the slice holds more than kInline candidates ------
  if (any_big) {
    __syncthreads();
    stamp();
    raster([&](int t, int c) {
      const uint32_t n = cf[c] & 0xffffu;
      if (n > (uint32_t)kInline) {
        const uint32_t k = atomicAdd(&cf[c], 0x10000u) >> 16;
        ids[word2[c] + k] = (uint16_t)t;
      }
    });
  }
  stamp();
}

// ------------------------------------------------------------------------------------------
// stream kernel
// ------------------------------------------------------------------------------------------
struct StreamParams {
  const float* points;
  void* out;
  unsigned char* ws;
  WsLayout L;
  int pts_stride, num_points, num_boxes, num_frames;
  int row_words;          // W
  int batch_pts;          // P: points per warp batch (32, or 1024 / W when W > 32)
  int batches_per_frame;  // ceil(N / P)
  int R, tb_base, tb_rem; // ranges of the frame-major batch list: range r has tb_base + (r < tb_rem) batches
  int slots;              // warps per range
  int vec4;
  int smem_prep;          // the CTA keeps the contract terms of one frame in shared memory
  int dynamic;            // CTA-local dynamic batch draws instead of static strides (experiment; default off)
  int grid_smem_words;    // GSM variant: words of the frame's cell grid copied to shared memory
  int rf;                 // frame-local variant: ranges per frame (R = rf * frames, tb_base / tb_rem per frame)
  unsigned long long* trace;  // profiling hook: 16 globaltimer stamps per warp (NULL = off)
};

// Index data is written by the prep kernel of the same PDL chain: plain (coherent, L1-cached)
// loads, never the non-coherent read-only path.
__device__ __forceinline__ uint32_t ld_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_u16(const uint16_t* p) {
  uint16_t v;
  asm volatile("ld.global.ca.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return (uint32_t)v;
}
__device__ __forceinline__ float4 ld_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void trace_stamp(const StreamParams& p, int k, int kWarps = 8) {
  if (p.trace && (threadIdx.x & 31) == 0 && k < 15) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[((size_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * 16 + k] = t;
    if (k == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      p.trace[((size_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * 16 + 15] = smid;
    }
  }
}

struct FrameCtx {
  float gx0, gy0, invx, invy, gp1;
  int Gp, f;
  const uint32_t* grid;
};

__device__ __forceinline__ FrameCtx load_frame(const StreamParams& p, int f) {
  const float4* h = reinterpret_cast<const float4*>(p.ws + p.L.hdr + (size_t)f * sizeof(FrameHdr));
  const float4 a = ld_f4(h);
  const uint32_t g = ld_u32(reinterpret_cast<const uint32_t*>(h) + 8);
  FrameCtx c;
  c.gx0 = a.x; c.gy0 = a.y; c.invx = a.z; c.invy = a.w;
  c.Gp = (int)g + 2;
  c.gp1 = (float)((int)g + 1);
  c.f = f;
  c.grid = reinterpret_cast<const uint32_t*>(p.ws + p.L.grid) + (size_t)f * p.L.gstride;
  return c;
}

__device__ __forceinline__ int cell_of(float x, float y, const FrameCtx& h) {
  return pcell(y, h.gy0, h.invy, h.gp1) * h.Gp + pcell(x, h.gx0, h.invx, h.gp1);
}

// The exact test of every candidate of cell word `w` for the point (x, y, z) of this lane;
// on_hit(t) for each enclosing box.  Sparse scenes take the two inline branches; the list
// loop serves dense scenes, where every lane of the warp is busy in it anyway.
// Contract terms of frame f: the CTA's shared-memory copy when it holds that frame, else global.
__device__ __forceinline__ const float4* prep_of(const StreamParams& p, int f, int f_smem, const float4* prep_smem) {
  return f == f_smem ? prep_smem : reinterpret_cast<const float4*>(p.ws + p.L.prep) + (size_t)f * 2 * p.num_boxes;
}

template <typename F>
__device__ __forceinline__ void for_each_hit(uint32_t w, const StreamParams& p, int f, const float4* prep, float x,
                                             float y, float z, F on_hit) {
  if (w == 0u) return;
  const int T = p.num_boxes;
  const uint32_t kind = w >> kKindShift;
  if (kind != 3u) {
    const uint32_t t0 = w & kIdMask;
    if (inside_box(x, y, z, prep[2 * t0], prep[2 * t0 + 1])) on_hit(t0);
    if (kind == 2u) {
      const uint32_t t1 = (w >> 15) & kIdMask;
      if (inside_box(x, y, z, prep[2 * t1], prep[2 * t1 + 1])) on_hit(t1);
    }
    return;
  }
  if (w == kCellAll) {
#pragma unroll 1
    for (int t = 0; t < T; ++t)
      if (inside_box(x, y, z, prep[2 * t], prep[2 * t + 1])) on_hit((uint32_t)t);
    return;
  }
  const uint16_t* ids = reinterpret_cast<const uint16_t*>(p.ws + p.L.ids) + (size_t)f * p.L.cap + (w & kListOffMask);
  uint32_t n = (w >> kListCntShift) & 0x3ffu;
  if (n == kCntLong) {
    n = ld_u16(ids);
    ++ids;
  }
#pragma unroll 1
  for (uint32_t j = 0; j < n; ++j) {
    const uint32_t t = ld_u16(ids + j);
    if (inside_box(x, y, z, prep[2 * t], prep[2 * t + 1])) on_hit(t);
  }
}

// WS > 0: row words known at compile time (the common shapes), 0: taken from the params.
// Work split: the frame-major list of P-point batches is cut into R contiguous ranges, range
// r is served by the CTAs with blockIdx % R == r (one SM's worth, sharing the frame's index in
// L1); warp `slot` of a range owns its batches slot, slot + slots, ... (static strides).
// Measured alternatives at the training shape: per-range L2 counters (+5 us:
// 80 same-address atomics per range at kernel start), CTA-local shared-memory counters (+2 us),
// a cp.async ring for the points (-0.5 us, +smem), two points per lane (same time), spreading
// a warp's batches over the frame (same time) — the kernel is instruction-issue bound with a
// tail of slow warps, not bandwidth bound.
// Software pipeline per warp: the points of batch n+2 are in flight, the cell word of batch
// n+1 is requested, batch n is tested and written.
// NT threads per CTA; GSM: one big CTA per SM that also keeps the cell grid of its frame in
// shared memory — a warp-wide gather of 32 scattered cell words costs ~32 L1 wavefronts from
// global memory but only a few bank-conflict replays from shared memory (the gathers were the
// hidden limiter of the L1 variant: same 18 us at every occupancy / block size).
template <int MODE, int WS, bool VEC4, int NT, bool GSM>
__global__ void __launch_bounds__(NT, GSM ? 1 : kOcc) pib_stream_kernel(const StreamParams p) {
  constexpr int kWarps = NT / 32;
  constexpr int kStreamThreads = NT;
  extern __shared__ __align__(16) uint32_t smem_all[];
  trace_stamp(p, 0, kWarps);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = WS > 0 ? WS : p.row_words;
  const int P = (MODE == kModePart || WS > 0) ? 32 : p.batch_pts;
  const int N = p.num_points, bpf = p.batches_per_frame, slots = p.slots;
  const int stage_words = P * W;
  // shared memory: [contract terms of one frame: 2 T float4, when they fit][per-warp stages]
  // [GSM: cell grid of one frame][contract terms of one frame][per-warp stages]
  const int grid_words = GSM ? p.grid_smem_words : 0;
  const int prep_words = p.smem_prep ? 8 * p.num_boxes : 0;
  const uint32_t* grid_smem = smem_all;
  const float4* prep_smem = reinterpret_cast<const float4*>(smem_all + grid_words);
  uint32_t* stage = smem_all + grid_words + prep_words + (size_t)warp * stage_words;

  const int r = blockIdx.x % p.R, slot = (int)(blockIdx.x / p.R) * kWarps + warp;
  const int nb = p.tb_base + (r < p.tb_rem ? 1 : 0);  // batches of this range
  const int g0 = r * p.tb_base + min(r, p.tb_rem);    // first batch of the range
  const int rf0 = g0 / bpf, rc0 = g0 - rf0 * bpf;     // its frame / batch inside the frame
  // the frame whose contract terms this CTA keeps in shared memory: that of its first batch
  const int f_cta = min(p.num_frames - 1, (g0 + (int)(blockIdx.x / p.R) * kWarps) / bpf);
  const int f_smem = p.smem_prep ? f_cta : -1;

  struct Batch { int f, c; };  // frame, batch inside the frame; f < 0: none
  auto decode = [&](int b) -> Batch {  // b: batch index inside the range, or < 0
    Batch t;
    t.f = -1; t.c = 0;
    if (b >= 0 && b < nb) {
      t.f = rf0; t.c = rc0 + b;
      while (t.c >= bpf) { t.c -= bpf; ++t.f; }
    }
    return t;
  };
  auto fetch = [&](const Batch& t, float& fx, float& fy, float& fz) {
    const int pt = t.c * P + lane;
    if (t.f >= 0 && lane < P && pt < N) {
      const size_t idx = (size_t)t.f * N + pt;
      if constexpr (VEC4) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p.points) + idx);
        fx = v.x; fy = v.y; fz = v.z;
      } else {
        const float* q = p.points + idx * p.pts_stride;
        fx = __ldcs(q); fy = __ldcs(q + 1); fz = __ldcs(q + 2);
      }
    }
  };
  // static first batch, requested before the index is ready
  Batch cur = decode(slot), nxt;
  float x = 0.f, y = 0.f, z = 0.f, x1 = 0.f, y1 = 0.f, z1 = 0.f;
  fetch(cur, x, y, z);

  trace_stamp(p, 1, kWarps);
  // everything below reads what the prep kernel wrote
  asm volatile("griddepcontrol.wait;" ::: "memory");
  trace_stamp(p, 2, kWarps);
  // batch draws: after its static batch a warp takes batches from its CTA's slab of the range
  // through a shared-memory counter (fast warps relieve slow ones; an L2 counter per range was
  // measured slower: 80 same-address atomics per range at kernel start cost ~5 us)
  __shared__ uint32_t s_draw;
  const bool dynamic = p.dynamic != 0 && cur.f >= 0;  // a warp without a static batch (tiny inputs) stays idle
  const int ncta = (int)((gridDim.x - 1 - (unsigned)r) / (unsigned)p.R) + 1;  // CTAs serving this range
  const int dyn_total = max(0, nb - slots);
  const int per_cta = (dyn_total + ncta - 1) / ncta;
  const int slab0 = slots + (int)(blockIdx.x / p.R) * per_cta, slab1 = min(nb, slab0 + per_cta);
  int static_next = slot + slots;  // static mode: strided batches
  auto draw = [&]() -> int {
    if (!dynamic) { const int b = static_next; static_next += slots; return b; }
    uint32_t k = 0;
    if (lane == 0) k = atomicAdd(&s_draw, 1u);
    const int b = slab0 + (int)__shfl_sync(0xffffffffu, k, 0);
    return b < slab1 ? b : -1;
  };
  if (threadIdx.x == 0) s_draw = 0u;
  if (p.smem_prep) {
    const float4* src = reinterpret_cast<const float4*>(p.ws + p.L.prep) + (size_t)f_smem * 2 * p.num_boxes;
    float4* dst = reinterpret_cast<float4*>(smem_all + grid_words);
    for (int k = threadIdx.x; k < 2 * p.num_boxes; k += kStreamThreads) dst[k] = ld_f4(src + k);
  }
  if constexpr (GSM) {
    const float4* src = reinterpret_cast<const float4*>(p.ws + p.L.grid + (size_t)f_cta * p.L.gstride * 4);
    float4* dst = reinterpret_cast<float4*>(smem_all);
    for (int k = threadIdx.x; k < (grid_words >> 2); k += kStreamThreads) dst[k] = ld_f4(src + k);
  }
  __syncthreads();

  FrameCtx fc;
  fc.f = -1;
  auto lookup = [&](const Batch& t, float lx, float ly) -> uint32_t {
    if (t.f < 0) return 0u;
    if (t.f != fc.f) fc = load_frame(p, t.f);
    if (lane >= P || t.c * P + lane >= N) return 0u;
    const int cell = cell_of(lx, ly, fc);
    if constexpr (GSM) {
      if (t.f == f_cta) return grid_smem[cell];
    }
    return ld_u32(fc.grid + cell);
  };
  nxt = decode(cur.f >= 0 ? draw() : -1);
  fetch(nxt, x1, y1, z1);
  uint32_t wn = lookup(cur, x, y);
  trace_stamp(p, 3, kWarps);
  int tk = 4;

#pragma unroll 1
  while (cur.f >= 0) {
    const int bf = cur.f, pt0 = cur.c * P;
    const int nvalid = min(P, N - pt0);
    const float cx = x, cy = y, cz = z;
    const uint32_t w = wn;
    // rotate the pipeline
    x = x1; y = y1; z = z1;
    cur = nxt;
    nxt = decode(cur.f >= 0 ? draw() : -1);  // batch n+2
    fetch(nxt, x1, y1, z1);
    wn = lookup(cur, x, y);

    if constexpr (MODE == kModePart) {
      uint32_t best = 0xffffffffu;
      for_each_hit(w, p, bf, prep_of(p, bf, f_smem, prep_smem), cx, cy, cz, [&](uint32_t t) { best = min(best, t); });
      if (lane < nvalid) __stcs(reinterpret_cast<int32_t*>(p.out) + (size_t)bf * N + pt0 + lane, (int32_t)best);
    } else {
      // zero the warp's stage (the linear image of its P rows), mark the hits, copy out
      if constexpr (WS > 0) {
#pragma unroll
        for (int k = 0; k < (32 * WS) / 128; ++k) reinterpret_cast<uint4*>(stage)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
        if ((32 * WS) % 128 != 0 && lane < (32 * WS % 128) / 4)
          reinterpret_cast<uint4*>(stage)[(32 * WS) / 128 * 32 + lane] = make_uint4(0, 0, 0, 0);
      } else {
#pragma unroll 1
        for (int k = lane; k < (stage_words >> 2); k += 32) reinterpret_cast<uint4*>(stage)[k] = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      for_each_hit(w, p, bf, prep_of(p, bf, f_smem, prep_smem), cx, cy, cz,
                   [&](uint32_t t) { stage[lane * W + (t >> 5)] |= 1u << (t & 31u); });
      __syncwarp();
      if constexpr (MODE == kModeBits) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.out) + ((size_t)bf * N + pt0) * W;
        const int nw = nvalid * W;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
          const int n4 = nw >> 2;
          if (WS > 0 && nvalid == 32 && (32 * WS) % 128 == 0) {
#pragma unroll
            for (int k = 0; k < (32 * WS) / 128; ++k)
              __stcs(reinterpret_cast<uint4*>(dst) + k * 32 + lane, reinterpret_cast<const uint4*>(stage)[k * 32 + lane]);
          } else {
#pragma unroll 1
            for (int k = lane; k < n4; k += 32)
              __stcs(reinterpret_cast<uint4*>(dst) + k, reinterpret_cast<const uint4*>(stage)[k]);
            for (int k = (n4 << 2) + lane; k < nw; k += 32) __stcs(dst + k, stage[k]);
          }
        } else {
          for (int k = lane; k < nw; k += 32) __stcs(dst + k, stage[k]);
        }
      } else {  // kModeAll: int32 [N, T] rows, lane l writes boxes 4l..4l+3 (+128 k)
        const int T = p.num_boxes;
        int32_t* dst = reinterpret_cast<int32_t*>(p.out) + ((size_t)bf * N + pt0) * T;
        if ((T & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
          for (int q = 0; q < nvalid; ++q) {
            for (int t4 = lane * 4; t4 < T; t4 += 128) {
              const uint32_t nib = stage[q * W + (t4 >> 5)] >> (t4 & 31);
              __stcs(reinterpret_cast<int4*>(dst + (size_t)q * T + t4),
                     make_int4(nib & 1u, (nib >> 1) & 1u, (nib >> 2) & 1u, (nib >> 3) & 1u));
            }
          }
        } else {
          for (int q = 0; q < nvalid; ++q)
            for (int t = lane; t < T; t += 32)
              __stcs(dst + (size_t)q * T + t, (int32_t)((stage[q * W + (t >> 5)] >> (t & 31)) & 1u));
        }
      }
      __syncwarp();
    }
    trace_stamp(p, tk++, kWarps);
  }
  trace_stamp(p, 14, kWarps);
}

// ------------------------------------------------------------------------------------------
// stream kernel, lean variant for the training shapes: bit-packed rows of 8 .. 32 words (129..1024
// boxes), 16-byte points.  Batch g of a frame is points[32 g .. 32 g + 31] and rows
// out[32 g ..] (16-byte aligned because rows are multiples of 16 bytes), and the per-batch
// bookkeeping of the general kernel (frame decode, alignment and tail paths: ~60 of its ~240
// warp instructions per batch) disappears.  The SM is instruction-issue bound in this kernel,
// so instructions are what is being saved.
// ------------------------------------------------------------------------------------------
// Frame-local ranges: every range lies inside one frame (R = rf ranges per frame x frames), so a
// CTA serves exactly one frame: one base pointer for its contract terms (no per-candidate select),
// the frame header and grid pointer are loop invariant and the frame-crossing bookkeeping of the
// loop disappears.  SP: the frame's terms are copied to shared memory first (candidate-heavy
// scenes); otherwise they are read from global memory through L1 (sparse scenes, 40 registers).
// FULL: N is a multiple of 32 (every batch has 32 points); otherwise the last batch of a frame is
// ragged: its loads are clamped to the frame's last point and its surplus rows are not stored.
template <int W, bool FULL, bool SP>
__global__ void __launch_bounds__(256, kOcc) pib_stream_frame_kernel(const StreamParams p) {
  constexpr int kW = 8;      // warps per CTA
  constexpr int kC = W / 4;  // 16-byte chunks of a row = store instructions per lane and batch
  extern __shared__ __align__(16) uint32_t smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = p.slots, T = p.num_boxes, N = p.num_points;
  const float4* prep_smem = reinterpret_cast<const float4*>(smem_all);
  uint32_t* stage = smem_all + (SP ? 8 * T : 0) + warp * (32 * W);

  const int r = blockIdx.x % p.R, cta = (int)(blockIdx.x / p.R);
  const int f = r / p.rf, rr = r - f * p.rf;
  const int g0 = rr * p.tb_base + min(rr, p.tb_rem);  // batches of this range, frame-local numbering
  const int gend = g0 + p.tb_base + (rr < p.tb_rem ? 1 : 0);
  int g = g0 + cta * kW + warp;                       // this warp's batches: g, g + stride, ...
  const float4* pts = reinterpret_cast<const float4*>(p.points) + (size_t)f * N;
  uint4* rows = reinterpret_cast<uint4*>(p.out) + (size_t)f * N * kC;
  auto load_pt = [&](int gb) { return __ldcs(pts + (FULL ? gb * 32 + lane : min(gb * 32 + lane, N - 1))); };
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
  if (g < gend) v0 = load_pt(g);
  if (g + stride < gend) v1 = load_pt(g + stride);

  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads what the prep kernel wrote
  const float4* prep_glob = reinterpret_cast<const float4*>(p.ws + p.L.prep) + (size_t)f * 2 * T;
  if constexpr (SP) {
    float4* dst = reinterpret_cast<float4*>(smem_all);
    for (int k = threadIdx.x; k < 2 * T; k += 256) dst[k] = ld_f4(prep_glob + k);
    __syncthreads();
  }
  if (g >= gend) return;

  const FrameCtx fc = load_frame(p, f);
  uint32_t wn = ld_u32(fc.grid + cell_of(v0.x, v0.y, fc));
#pragma unroll 1
  for (;;) {
    const float4 v = v0;
    const uint32_t w = wn;
    const int gn = g + stride;
    // rotate: points of batch g + 2 stride, cell word of batch g + stride
    v0 = v1;
    if (gn + stride < gend) v1 = load_pt(gn + stride);
    if (gn < gend) wn = ld_u32(fc.grid + cell_of(v0.x, v0.y, fc));
#pragma unroll
    for (int k = 0; k < kC; ++k) reinterpret_cast<uint4*>(stage)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    for_each_hit(w, p, f, SP ? prep_smem : prep_glob, v.x, v.y, v.z,
                 [&](uint32_t t) { stage[lane * W + (t >> 5)] |= 1u << (t & 31u); });
    __syncwarp();
    uint4* dst = rows + (size_t)g * (32 * kC) + lane;
    const int lim = FULL ? 32 * kC : min(32, N - g * 32) * kC;  // 16-byte chunks of the batch's rows
#pragma unroll
    for (int k = 0; k < kC; ++k)
      if (FULL || k * 32 + lane < lim) __stcs(dst + k * 32, reinterpret_cast<const uint4*>(stage)[k * 32 + lane]);
    __syncwarp();
    if (gn >= gend) break;
    g = gn;
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

template <int MODE, int WS, bool VEC4, int NT = kStreamThreads, bool GSM = false>
int launch_stream(const StreamParams& p, int grid, size_t smem, cudaStream_t st) {
  static int configured[64];
  int dev = 0;
  GGA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && (int)smem > configured[dev] && smem > 48 * 1024) {
    GGA_CHECK_CUDA(
        cudaFuncSetAttribute(pib_stream_kernel<MODE, WS, VEC4, NT, GSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = (int)smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GGA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pib_stream_kernel<MODE, WS, VEC4, NT, GSM>, p));
  return GGA_OK;
}

int run_pib(int mode, const float* points, int pts_stride, const float* boxes, void* out, int B, int num_points,
            int num_boxes, void* workspace, size_t workspace_bytes, void* stream) {
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
  if (num_boxes > kMaxBoxes) {
    gga_set_error("num_boxes=%d exceeds the %d boxes per frame this build indexes", num_boxes, kMaxBoxes);
    return GGA_ERR_UNSUPPORTED;
  }
  const int Gmax = pick_gmax(num_points, num_boxes);
  const WsLayout L = ws_layout(B, num_boxes, Gmax);
  GGA_REQUIRE(workspace != nullptr, "null workspace (size it with gga_pib_workspace_bytes, zero it once)");
  GGA_REQUIRE(workspace_bytes >= L.total, "workspace too small: %zu bytes given, %zu needed", workspace_bytes,
              L.total);
  GGA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");

  const int nsm = gga_sm_count();
  const int occ = g_tune_occ > 0 ? (g_tune_occ > kOcc ? kOcc : g_tune_occ) : kOcc;

  StreamParams sp;
  sp.points = points; sp.out = out; sp.ws = static_cast<unsigned char*>(workspace); sp.L = L;
  sp.trace = g_trace;
  sp.pts_stride = pts_stride; sp.num_points = num_points; sp.num_boxes = num_boxes; sp.num_frames = B;
  sp.row_words = gga_pib_row_words(num_boxes);
  sp.batch_pts = sp.row_words <= 32 ? 32 : (1024 / sp.row_words < 1 ? 1 : 1024 / sp.row_words);
  if (mode == kModePart) sp.batch_pts = 32;
  sp.batches_per_frame = (num_points + sp.batch_pts - 1) / sp.batch_pts;
  const long long tb = (long long)sp.batches_per_frame * B;
  GGA_REQUIRE(tb < (1ll << 30), "too many point batches in one call");
  sp.slots = occ * kWarps;
  long long R = (tb + sp.slots - 1) / sp.slots;
  // More ranges than SMs: the grid is then 2-3x what is resident at once (5 CTAs per SM) and the
  // hardware hands freed slots to waiting CTAs, which evens out the spread between cheap and
  // candidate-heavy batches (measured on B200: c2 22.9 -> 20.7 us, c3 84 -> 70 us, c5 87 -> 82 us).
  const double bpw = (double)tb / ((double)nsm * sp.slots);  // batches per warp with one range per SM
  const int range_mult = (sp.row_words >= 16 || bpw >= 7.5) ? 3 : 2;
  if (R > (long long)nsm * range_mult) R = (long long)nsm * range_mult;
  if (R < 1) R = 1;
  sp.R = (int)R;
  sp.tb_base = (int)(tb / R);
  sp.tb_rem = (int)(tb % R);
  sp.dynamic = 0;
  sp.rf = 0;
  GGA_REQUIRE(sp.R <= kMaxRanges, "too many ranges");
  sp.vec4 = (pts_stride == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0) ? 1 : 0;
  const int grid = (int)R * occ;
  sp.smem_prep = num_boxes <= 512 ? 1 : 0;
  const size_t smem = (mode == kModePart ? 0 : (size_t)sp.row_words * sp.batch_pts * 4 * kWarps) +
                      (sp.smem_prep ? (size_t)num_boxes * 32 : 0);
  GGA_REQUIRE(smem <= 200 * 1024, "row too wide for the stage");

  // prep: S slices per frame
  int S = nsm / B;
  if (S < 1) S = 1;
  if (S > Gmax + 2) S = Gmax + 2;
  {
    const int gp = Gmax + 2;
    const int min_s = (gp * gp + kMaxSliceCells - 1) / kMaxSliceCells + 1;  // keeps Gmax reachable
    if (S < min_s && (long long)B * min_s <= 4ll * nsm) S = min_s;
  }
  PrepParams pp;
  pp.boxes = boxes; pp.ws = sp.ws; pp.L = L; pp.T = num_boxes;
  pp.ncache = num_boxes < 2048 ? num_boxes : 2048;
  pp.trace = g_trace_prep;
  {
    int Gcap = Gmax;  // slices are row bands of the padded grid: ceil((G+2)/S) rows of G+2 cells
    while (Gcap > 1 && ((Gcap + 2 + S - 1) / S) * (Gcap + 2) > kMaxSliceCells) --Gcap;
    pp.Gcap = Gcap;
    pp.max_slice_cells = ((Gcap + 2 + S - 1) / S) * (Gcap + 2);
    if (pp.max_slice_cells < 9) pp.max_slice_cells = 9;  // G = 1: 3 x 3
    if (pp.max_slice_cells > kMaxSliceCells) pp.max_slice_cells = kMaxSliceCells;
  }
  const size_t prep_smem = (size_t)3 * (kMaxSliceCells + 4) * 4 + align_up((size_t)num_boxes * 12, 16) +
                           (size_t)pp.ncache * kBoxWords * 4;
  {
    static int configured[64];
    int dev = 0;
    GGA_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && (int)prep_smem > configured[dev] && prep_smem > 48 * 1024) {
      GGA_CHECK_CUDA(
          cudaFuncSetAttribute(pib_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
      configured[dev] = (int)prep_smem;
    }
  }
  if (g_tune_phase != 2 && g_tune_phase != 3) {
    pp.S = S;
  pib_prep_kernel<<<dim3(S + (num_boxes + kPrepThreads - 1) / kPrepThreads, B), kPrepThreads, prep_smem, st>>>(pp);
    GGA_CHECK_CUDA(cudaGetLastError());
  }
  if (g_tune_phase == 1) return GGA_OK;

  // lean variant: bit-packed rows of 8 / 16 / 24 / 32 words (129..1024 boxes), 16-byte points
  if (mode == kModeBits && sp.vec4 && !g_tune_nofast && kStreamThreads == 256 && (sp.row_words & 7) == 0 &&
      sp.row_words <= 32) {
    // W = 8 (<= 256 boxes, sparse candidates): contract terms straight from global memory through
    // L1 — a CTA touches a few dozen boxes, copying all of them to shared memory per CTA costs more
    // than it saves (measured: 19.7 -> 18.4 us at c2), and the kernel then fits 40 registers.
    // W = 16 (<= 512 boxes, candidate-heavy scenes): shared-memory copy (c3: 59.5 vs 65.6 us).
    // W = 24 / 32: global memory again (32 KB of terms per CTA would halve the occupancy); c5 82 -> 74 us.
    // (6 CTAs per SM with the 40-register variant: no gain.)
    const bool use_sp = sp.row_words == 16;
    const int occ_l = kOcc;
    StreamParams fp = sp;
    fp.slots = occ_l * 8;
    long long Rl = (tb + fp.slots - 1) / fp.slots;
    if (Rl > (long long)nsm * range_mult) Rl = (long long)nsm * range_mult;
    long long rf = (Rl + B - 1) / B;   // ranges per frame
    if (rf > sp.batches_per_frame) rf = sp.batches_per_frame;
    if (rf < 1) rf = 1;
    fp.rf = (int)rf;
    fp.R = (int)rf * B;
    fp.tb_base = (int)(sp.batches_per_frame / rf);
    fp.tb_rem = (int)(sp.batches_per_frame % rf);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fp.R * occ_l);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (use_sp ? (size_t)num_boxes * 32 : 0) + (size_t)sp.row_words * 32 * 4 * 8;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const bool full = (num_points & 31) == 0;
#define GGA_LEAN(W_, FULL_, SP_) GGA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pib_stream_frame_kernel<W_, FULL_, SP_>, fp))
    if (sp.row_words == 8) {
      if (full) GGA_LEAN(8, true, false); else GGA_LEAN(8, false, false);
    } else if (sp.row_words == 24) {
      if (full) GGA_LEAN(24, true, false); else GGA_LEAN(24, false, false);
    } else if (sp.row_words == 32) {
      if (full) GGA_LEAN(32, true, false); else GGA_LEAN(32, false, false);
    } else {
      if (full) GGA_LEAN(16, true, true); else GGA_LEAN(16, false, true);
    }
#undef GGA_LEAN
    return GGA_OK;
  }
  {
    // one 1024-thread CTA per SM with the frame's cell grid in shared memory, when it all fits
    constexpr int kBig = 1024;
    const size_t gbytes = L.gstride * 4;
    const size_t smem_big = gbytes + (sp.smem_prep ? (size_t)num_boxes * 32 : 0) +
                            (size_t)sp.row_words * 32 * 4 * (kBig / 32);
    if (mode == kModeBits && sp.vec4 && sp.smem_prep && sp.batch_pts == 32 && smem_big <= 200 * 1024 && g_tune_gsm) {
      StreamParams bp = sp;
      bp.grid_smem_words = (int)L.gstride;
      bp.slots = kBig / 32;
      long long Rb = (tb + bp.slots - 1) / bp.slots;
      if (Rb > nsm) Rb = nsm;
      if (Rb < 1) Rb = 1;
      bp.R = (int)Rb;
      bp.tb_base = (int)(tb / Rb);
      bp.tb_rem = (int)(tb % Rb);
      if (sp.row_words == 8) return launch_stream<kModeBits, 8, true, kBig, true>(bp, (int)Rb, smem_big, st);
      return launch_stream<kModeBits, 0, true, kBig, true>(bp, (int)Rb, smem_big, st);
    }
  }
  sp.grid_smem_words = 0;
  if (mode == kModeBits) {
    if (sp.vec4) {
      if (sp.row_words == 8) return launch_stream<kModeBits, 8, true>(sp, grid, smem, st);
      if (sp.row_words == 2) return launch_stream<kModeBits, 2, true>(sp, grid, smem, st);
      return launch_stream<kModeBits, 0, true>(sp, grid, smem, st);
    }
    return launch_stream<kModeBits, 0, false>(sp, grid, smem, st);
  }
  if (mode == kModeAll) return launch_stream<kModeAll, 0, false>(sp, grid, smem, st);
  return launch_stream<kModePart, 0, false>(sp, grid, smem, st);
}

}  // namespace

extern "C" int gga_pib_row_words(int num_boxes) {
  if (num_boxes <= 0) return 0;
  if (num_boxes <= 32) return 1;
  if (num_boxes <= 64) return 2;
  if (num_boxes <= 128) return 4;
  return 8 * ((num_boxes + 255) / 256);
}

extern "C" int gga_pib_set_tuning(int grid_cells, int ctas_per_sm) {
  g_tune_grid = grid_cells;
  g_tune_nogsm = ctas_per_sm < 0 ? 1 : 0;   // negative: never the shared-memory-grid variant of the stream kernel
  g_tune_nofast = ctas_per_sm != 0 ? 1 : 0;  // any explicit value: never the lean training-shape variant either
  g_tune_gsm = (ctas_per_sm > 0 && ctas_per_sm >= 100) ? 1 : 0;  // +100: the shared-memory-grid variant (measured: no gain)
  g_tune_occ %= 100;
  g_tune_occ = ctas_per_sm < 0 ? -ctas_per_sm : ctas_per_sm;
  return GGA_OK;
}

/* profiling hook: device buffer of 16 x uint64 per stream-kernel warp (globaltimer ns, last word = smid) */
extern "C" int gga_test_pib_trace(void* buf) {
  g_trace = static_cast<unsigned long long*>(buf);
  return GGA_OK;
}
extern "C" int gga_test_pib_trace_prep(void* buf) {
  g_trace_prep = static_cast<unsigned long long*>(buf);
  return GGA_OK;
}

/* profiling hook: 0 = both kernels, 1 = index build only, 2 = streaming only (reuses the index in the workspace) */
extern "C" int gga_test_pib_phase(int phase) {
  g_tune_phase = phase;
  return GGA_OK;
}

extern "C" size_t gga_pib_workspace_bytes(int B, int num_points, int num_boxes) {
  if (B <= 0 || num_points <= 0 || num_boxes <= 0) return 256;
  if (num_boxes > kMaxBoxes) num_boxes = kMaxBoxes;
  return ws_layout(B, num_boxes, pick_gmax(num_points, num_boxes)).total;
}

extern "C" int gga_pib_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  GGA_REQUIRE(workspace != nullptr || workspace_bytes == 0, "null workspace");
  if (workspace_bytes == 0) return GGA_OK;
  GGA_CHECK_CUDA(cudaMemsetAsync(workspace, 0, workspace_bytes, gga_stream(stream)));
  return GGA_OK;
}

extern "C" int gga_points_in_boxes_bits(const float* points, int pts_stride, const float* boxes, uint32_t* bits,
                                        int B, int num_points, int num_boxes, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return run_pib(kModeBits, points, pts_stride, boxes, bits, B, num_points, num_boxes, workspace,
                 workspace_bytes, stream);
}

extern "C" int gga_points_in_boxes_all(const float* points, int pts_stride, const float* boxes, int32_t* out,
                                       int B, int num_points, int num_boxes, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  return run_pib(kModeAll, points, pts_stride, boxes, out, B, num_points, num_boxes, workspace, workspace_bytes,
                 stream);
}

extern "C" int gga_points_in_boxes_part(const float* points, int pts_stride, const float* boxes, int32_t* out,
                                        int B, int num_points, int num_boxes, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return run_pib(kModePart, points, pts_stride, boxes, out, B, num_points, num_boxes, workspace,
                 workspace_bytes, stream);
}

extern "C" int gga_points_in_boxes_all_host(const float* points, int pts_stride, const float* boxes,
                                            int32_t* out, int B, int num_points, int num_boxes) {
  GGA_REQUIRE(B >= 0 && num_points >= 0 && num_boxes >= 0, "negative size");
  if (B == 0 || num_points == 0 || num_boxes == 0) return GGA_OK;
  GGA_REQUIRE(points && boxes && out, "null pointer");
  const size_t pb = (size_t)B * num_points * pts_stride * sizeof(float);
  const size_t bb = (size_t)B * num_boxes * 7 * sizeof(float);
  const size_t ob = (size_t)B * num_points * num_boxes * sizeof(int32_t);
  const size_t wb = gga_pib_workspace_bytes(B, num_points, num_boxes);
  float *dp = nullptr, *db = nullptr;
  int32_t* dout = nullptr;
  void* dws = nullptr;
  cudaStream_t st;
  GGA_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = GGA_OK;
  cudaError_t e = cudaMallocAsync(&dp, pb, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&db, bb, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&dout, ob, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&dws, wb, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dws, 0, wb, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dp, points, pb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, boxes, bb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = run_pib(kModeAll, dp, pts_stride, db, dout, B, num_points, num_boxes, dws, wb, st);
    if (rc == GGA_OK) e = cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, st);
  }
  if (dp) cudaFreeAsync(dp, st);
  if (db) cudaFreeAsync(db, st);
  if (dout) cudaFreeAsync(dout, st);
  if (dws) cudaFreeAsync(dws, st);
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
