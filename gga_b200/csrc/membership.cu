// Part 1 — point -> 3D-box membership for sm_100a.
//
// Contract (bit-exact with the CPU op): SURVEY.md Appendix A.1; replaces
// mmcv.ops.points_in_boxes_{all,part,cpu} (re-exported at
// /root/reference/mmdet3d/ops/__init__.py:12-13, called from
// /root/reference/mmdet3d/core/bbox/structures/base_box3d.py:534,566).
//
// Design.  The brute-force test is FP32-issue bound (14 instr x N x T); the useful
// traffic is 16 B in and 4*W B out per point.  ONE persistent kernel, no global scratch:
//
//  * a CTA serves the 256-point tiles r, r + rf, ... of ONE frame.  It first builds that frame's box
//    index in its own shared memory (while its first point tile is already in flight):
//      - the exact per-box contract terms (cos/sin of -rz through include/gga_detmath.h),
//      - the conservative BEV rectangle of every box (box_rect) and, from their extent, ONE clamped
//        monotone bin function shared by rectangles and points (bin_of),
//      - at most 256 boxes: a CELL TABLE (cells of 2 x 2 bins, up to 7 one-byte box ids + count per
//        cell) filled with shared-memory atomics by all threads of the CTA,
//      - more boxes, or a scene so dense that a cell overflows: two AXIS MASK tables — for each bin
//        along x (and along y) a row of W words whose bit t says "the rectangle of box t overlaps this
//        bin"; maskx[bx] & masky[by] = the candidate boxes AS A BIT ROW in the output layout.
//  * points enter through a ring of TMA bulk copies (full / empty mbarrier per slot); per point: two
//    bin numbers, one cell entry (or two mask rows), the exact test only for the candidates.
//  * rows are staged per warp in shared memory as the linear image of the output (chunk order rotated
//    per lane: conflict free) and leave as ONE bulk copy per 32-point batch.
//  * more than 1024 boxes: the box list is served in chunks of 1024 (32 row words) per sweep.
//  * launched with programmatic dependent launch: table reset and barrier set-up run before
//    griddepcontrol.wait, i.e. under the previous kernel's tail.
//
// Culling never changes the result; every candidate is decided by outside_z() / inside_xy() below.
#include <float.h>

#include "../../include/gga_detmath.h"
#include "common.cuh"

namespace {

constexpr int kBins = 128;         // bins per axis (126 inner bins tile the boxes' extent, 2 border bins)
constexpr int kChunkBoxes = 1024;  // boxes indexed per sweep = 32 row words
constexpr int kTileBatches = 8;    // point tile of the shared-memory ring: 8 batches = 256 points
constexpr int kMaxSlots = 56;      // ring slots (tiles) a CTA can hold
// Cell table of the sparse path (at most 256 boxes): kCellsSide^2 cells of 2 x 2 bins, each an 8-byte
// entry with the ids (one byte each) of up to 7 boxes whose rectangle touches the cell and their count.
// A slot is claimed with an atomicAdd on the count; shared-memory atomics that return a value have
// several hundred cycles of latency (measured), so ALL threads of the CTA share the insertions
// (4 threads per box at 256 boxes, four cells in flight per thread).
constexpr int kCellsSide = kBins / 2;
constexpr int kCellIds = 7;
constexpr int kZBins = 32;       // bins along z of the general path's third mask (30 inner, 2 border)
constexpr int kWideCells = 256;  // a box covering more cells is kept in a short list tested for every point
constexpr int kMaxWide = 8;
__host__ __device__ __forceinline__ int cell_of_bin(int b) { return b >> 1; }

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
__device__ __forceinline__ bool inside_xy(float x, float y, const float4 a, const float4 r) {
  const float sx = __fsub_rn(x, a.x), sy = __fsub_rn(y, a.y);
  const float lx = __fadd_rn(__fmul_rn(sx, r.x), __fmul_rn(sy, -r.y));
  const float ly = __fadd_rn(__fmul_rn(sx, r.y), __fmul_rn(sy, r.x));
  return (lx > -r.z) & (lx < r.z) & (ly > -r.w) & (ly < r.w);
}
__device__ __forceinline__ bool outside_z(float z, const float4 a) {
  return fabsf(__fsub_rn(z, a.z)) > a.w;  // NaN z passes, like the CPU op
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

// Conservative z slab of a box for the z mask of the general path.  The contract's test is
// |fl(pz - cz)| <= RD(dz/2) with cz = fl(z + dz/2 in double); zc below is within 2 ulp of that cz and
// the slack covers it and the rounding of the subtraction.  kind: 0 = no finite point can pass (the
// last z bin, where NaN z lands, lists every box anyway), 1 = finite slab, 2 = every z bin.
__device__ __forceinline__ int box_zslab(float z, float dz, float& lo, float& hi) {
  const float h = 0.5f * dz, zc = z + h;
  if (!(h == h) || !isfinite(zc)) return 2;   // NaN size or non-finite centre: cannot be ordered
  if (h < 0.f) return 0;
  if (!isfinite(h)) return 2;
  const float slack = __fadd_ru(__fmul_ru(__fadd_ru(fabsf(zc), h), 9.5367431640625e-7f), 1e-37f);
  lo = __fsub_rd(__fsub_rd(zc, h), slack);
  hi = __fadd_ru(__fadd_ru(zc, h), slack);
  if (!isfinite(lo) || !isfinite(hi)) return 2;
  return 1;
}

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct SweepParams {
  const float* points;  // [B, N, pts_stride]
  const float* boxes;   // [B, T, 7]
  void* out;            // bits: uint32 [B, N, W]; all: int32 [B, N, T]; part: int32 [B, N]
  int pts_stride, N, T, B;
  int W;           // words per row of the bit-packed layout (also sizes the shared-memory tables)
  int bpf;         // 32-point batches per frame
  int nchunks;     // sweeps over the box list (1024 boxes each)
  int ring_slots;  // point tiles the shared-memory ring holds (0: points are loaded directly)
  unsigned long long* trace;  // GGA_PROFILING builds: 16 globaltimer stamps per warp (NULL = off)
  int variant;                // GGA_PROFILING builds: experiment switches (0 in the product)
};

// mask rows of 8+ words are padded by one 16-byte chunk: an odd chunk stride spreads the random
// row gathers of a warp over all 8 bank groups (measured: 12.6 -> 9.0 cycles per LDS.128)
__host__ __device__ __forceinline__ int mask_stride(int wc) { return wc >= 8 ? wc + 4 : wc; }

// Bin of a coordinate: one FMA (a single rounding of an affine function with non-negative slope
// is monotone), two clamps, truncation.  Rectangles and points go through the same function, so
// rect.lo <= p <= rect.hi implies bin(rect.lo) <= bin(p) <= bin(rect.hi).  Bins 0 and kBins-1 catch
// everything outside the boxes' extent; NaN maps to the last bin (fminf returns the non-NaN operand).
__device__ __forceinline__ int bin_of(float v, float inv, float off) {
  return (int)fmaxf(fminf(__fmaf_rn(v, inv, off), (float)(kBins - 1)), 0.f);
}
__device__ __forceinline__ int zbin_of(float v, float inv, float off) {  // same construction, kZBins bins
  return (int)fmaxf(fminf(__fmaf_rn(v, inv, off), (float)(kZBins - 1)), 0.f);
}

// ---- bulk asynchronous copies (TMA, 1-D) and their mbarriers -------------------------------------
// Points enter through a shared-memory ring filled by cp.async.bulk (one elected lane requests 4 KB
// tiles; the whole share of a CTA is in flight from the first microsecond: the sweep of a
// training-shape launch lasts only a few DRAM latencies, so register prefetch of 2-3 batches per
// warp left it latency bound) and finished rows leave through cp.async.bulk as well, so they cross
// the load/store unit once (the STS that built them) instead of three times (STS, LDS, STG).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {  // bar_addr: shared-space address
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MODE: output layout.  WC: row words of a sweep known at compile time (1, 2, 4, 8: at most 256
// boxes, the sparse path) or 0 = a multiple of 8 taken from the params (up to 32).  NT: threads per
// CTA.  VEC4: 16-byte points.
template <int MODE, int WC, int NT, bool VEC4>
__global__ void __launch_bounds__(NT, 1) pib_sweep_kernel(const SweepParams p) {
  constexpr int kWarps = NT / 32;
  constexpr int KB = (kChunkBoxes + NT - 1) / NT;     // box jobs per thread in the index build
  constexpr bool kWide = (WC == 0 || WC >= 4);        // rows of whole 16-byte chunks
  constexpr bool kBulk = MODE == kModeBits && kWide;  // rows can leave through the bulk copy engine
  constexpr bool kCells = WC != 0;                    // <= 256 boxes: candidates come from the cell table
  constexpr int kStages = kBulk && WC != 0 ? 2 : 1;   // stage buffers per warp (narrow rows: double buffered)
  constexpr int R = kWarps / kTileBatches;            // point tiles consumed at a time (8 warps per tile)
  // general path, 1024-thread CTAs (bit rows of 16 / 24 words): a third mask along z.  Dense indoor
  // scenes lose ~half of their candidates to it (c3 56 -> 48 us); with 32-word rows the extra 128-byte
  // gather per point costs more than the candidates it removes (c5 87 -> 113 us), so not there.
  constexpr bool kZ = WC == 0 && NT == 1024;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_wext[kWarps][6];  // per-warp extent of the finite box rectangles (and z slabs)
  __shared__ float s_bin[6];              // bin function of the sweep: invx, offx, invy, offy, invz, offz
  __shared__ int s_flags[2];              // [0] some cell holds more than kCellIds boxes (axis masks needed), [1] cells unusable
  __shared__ int s_nwide;
  __shared__ int s_wide[kMaxWide];
  __shared__ uint32_t s_rect[kCells ? 256 : 1];  // bin rectangle of every box (sparse path), 4 x 8 bits
  __shared__ __align__(8) unsigned long long s_full[kMaxSlots], s_empty[kMaxSlots];

  // Programmatic dependent launch: the next kernel of the stream may become resident as this one
  // drains.  Everything that touches no global memory (parameters, addresses, zeroed tables,
  // barrier set-up) runs BEFORE the wait and so overlaps the previous kernel's tail; nothing is
  // read or written in global memory before the previous kernel of the stream has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rf = gridDim.x, rr = blockIdx.x;
  const int bpf = p.bpf, N = p.N;
  const int tr = warp >> 3, tj = warp & 7;  // this warp serves batch tj of the tiles tr, tr + R, ...
  int tk = 0;
  auto stamp = [&]() {  // GGA_PROFILING: 16 globaltimer stamps per warp
#ifdef GGA_PROFILING
    if (p.trace && lane == 0 && tk < 16) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p.trace[(((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kWarps + warp) * 16 + tk] = t;
    }
#endif
    ++tk;
  };

  const int wcap = WC > 0 ? WC : min(p.W, 32);
  const int tcap = min(p.T, kChunkBoxes);
  const int mcap = mask_stride(wcap);
  float4* terms = reinterpret_cast<float4*>(smem_raw);           // [2 tcap]
  uint32_t* mx = reinterpret_cast<uint32_t*>(terms + 2 * tcap);  // [kBins][mask_stride(wc)]
  uint32_t* my = mx + kBins * mcap;
  uint32_t* mz = my + kBins * mcap;                              // [kZBins][mask_stride(wc)] (general path only)
  constexpr int kMaskRows = 2 * kBins + (kZ ? kZBins : 0);
  // per-warp stage: the linear image of the 32 rows of a batch; two buffers when the rows leave
  // through the bulk copy engine (one is being read by it while the next batch is built)
  uint32_t* stage_all = mx + kMaskRows * mcap;
  uint32_t* stage_base = stage_all + warp * (kStages * 32 * wcap);
  uint32_t* cells = stage_all + kWarps * (kStages * 32 * wcap);  // [kCellsSide^2][2] (sparse path only)
  unsigned char* ring = reinterpret_cast<unsigned char*>(cells + (kCells ? kCellsSide * kCellsSide * 2 : 0));
  const uint32_t tile_bytes = 1024u * (uint32_t)p.pts_stride;  // 256 points
  const int S = p.ring_slots / R;                              // ring slots of this warp group
  auto reset_tables = [&](bool again) {
    for (int i = tid; i < (kMaskRows * mcap) >> 2; i += NT) reinterpret_cast<uint4*>(mx)[i] = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (kCells) {
      for (int i = tid; i < (kCellsSide * kCellsSide * 2) >> 2; i += NT) reinterpret_cast<uint4*>(cells)[i] = make_uint4(0u, 0u, 0u, 0u);
      if (tid == 0) { s_flags[0] = 0; s_flags[1] = 0; s_nwide = 0; }
    }
    if (tid < p.ring_slots) {
      if (again) {
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_full[tid])) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_empty[tid])) : "memory");
      }
      mbar_init(&s_full[tid], 1u);
      mbar_init(&s_empty[tid], (uint32_t)kTileBatches);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  };
  reset_tables(false);
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  stamp();

  bool first = true;
  for (int f = blockIdx.y; f < p.B; f += gridDim.y) {
    const float* __restrict__ pts = p.points + (size_t)f * N * p.pts_stride;
    auto load_pt = [&](int gb) -> float4 {  // lane's point of batch gb (clamped to the frame's last point)
      const int i = min(gb * 32 + lane, N - 1);
      if constexpr (VEC4) {
        return __ldcs(reinterpret_cast<const float4*>(pts) + i);
      } else {
        const float* q = pts + (size_t)i * p.pts_stride;
        return make_float4(__ldcs(q), __ldcs(q + 1), __ldcs(q + 2), 0.f);
      }
    };

    for (int ch = 0; ch < p.nchunks; ++ch) {
      const int t0 = ch * kChunkBoxes;
      const int Tc = min(kChunkBoxes, p.T - t0);
      const int wc = WC > 0 ? WC : min(32, p.W - 32 * ch);
      const int ms = mask_stride(wc);
      if (!first) {
        __syncthreads();  // the previous sweep is done with the tables and the ring
        reset_tables(true);
        __syncthreads();
      }
      first = false;

      // ---- point tiles: CTA r of a frame serves the 256-point tiles r, r + rf, ...; a group of 8
      //      warps consumes one tile (a batch each), the group's warp 0 / lane 0 requests its tiles
      const int ntiles = (bpf + kTileBatches - 1) / kTileBatches;
      const bool ring_ok = S > 0 && (reinterpret_cast<uintptr_t>(pts) & 15u) == 0;
      const int n_full = ring_ok ? (N >> 8) : 0;  // whole tiles go through the ring, a ragged last one is loaded directly
      auto request_tile = [&](int k, int slot) {  // k-th tile of this warp group into ring slot `slot`
        const int ft = rr + rf * (tr + R * k);
        if (ft < n_full) {
          mbar_expect_tx(&s_full[slot], tile_bytes);
          bulk_load(ring + (size_t)slot * tile_bytes, reinterpret_cast<const unsigned char*>(pts) + (size_t)ft * tile_bytes,
                    tile_bytes, &s_full[slot]);
        }
      };
      // The first tile of every warp group is requested right away; the rest of the ring once the
      // boxes have arrived (they are the head of the index build's dependent chain and would
      // otherwise queue behind ~100 KB of point tiles per SM in the memory system).
      auto request_rest = [&]() {
#ifdef GGA_PROFILING
        if (p.variant & 512) return;
#endif
        if (tj == 0 && lane == 0)
          for (int k = 1; k < S; ++k) request_tile(k, tr + R * k);
      };
      if (tj == 0 && lane == 0 && S > 0) request_tile(0, tr);

      // ---- index build -------------------------------------------------------------------
      // Box job j (rectangle, extent) belongs to thread j mod NT.  The exact contract terms of box j
      // (double precision, ~0.7 us) are computed by thread (j + eoff) mod NT: when the boxes occupy
      // at most half of the CTA, other warps do that concurrently.
      const int nbw = (min(Tc, NT) + 31) >> 5;  // warps holding box jobs
      const int eoff = 2 * nbw <= kWarps ? 32 * nbw : 0;
      const float* __restrict__ fb = p.boxes + ((size_t)f * p.T + t0) * 7;
      const bool box_warp = warp < nbw, terms_warp = eoff == 0 ? box_warp : (warp >= nbw && warp < 2 * nbw);
      float bq[KB][7];
      if (box_warp || terms_warp) {
        const int e = eoff != 0 && !box_warp ? tid - eoff : tid;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const float* b = fb + (size_t)min(e + k * NT, Tc - 1) * 7;
#pragma unroll
          for (int j = 0; j < 7; ++j) bq[k][j] = __ldg(b + j);
        }
      }
      stamp();
      float invx = 0.f, invy = 0.f, offx = 1.f, offy = 1.f, invz = 0.f, offz = 1.f;
      auto exact_terms = [&]() {
        const int e = eoff != 0 ? tid - eoff : tid;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int t = e + k * NT;
          if (t < Tc) {
            const BoxPrep q = prep_box(bq[k]);
            terms[2 * t] = make_float4(q.cx, q.cy, q.cz, q.hz);
            terms[2 * t + 1] = make_float4(q.cosa, q.sina, q.hx, q.hy);
          }
        }
      };
      int bx0[KB], bx1[KB], by0[KB], by1[KB], kind[KB];
      int bz0[KB], bz1[KB];  // z bins of the box's slab (general path)
      if (box_warp) {
        // the box warps: rectangles and extent; they synchronise among themselves on named barrier 1
        float rx0[KB], rx1[KB], ry0[KB], ry1[KB], rz0[KB], rz1[KB];
        int zkind[KB];
        uint32_t mnx = 0xffffffffu, mny = 0xffffffffu, mxx = 0u, mxy = 0u, mnz = 0xffffffffu, mxz = 0u;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int t = tid + k * NT;
          kind[k] = 0;
          rx0[k] = rx1[k] = ry0[k] = ry1[k] = 0.f;
          if (t < Tc) {
            // fp32 trig is enough for a CONSERVATIVE rectangle: the 2^-13 inflation of box_rect dwarfs
            // its error (__sincosf: 2^-21.4 absolute inside [-pi, pi]; sincosf: 2 ulp elsewhere)
            float sn, cs;
            const float rz = bq[k][6];
            if (fabsf(rz) <= 3.2f) __sincosf(-rz, &sn, &cs);
            else sincosf(-rz, &sn, &cs);
            kind[k] = box_rect(bq[k][0], bq[k][1], cs, sn, __fmul_ru(bq[k][3], 0.5f), __fmul_ru(bq[k][4], 0.5f), rx0[k],
                               rx1[k], ry0[k], ry1[k]);
            if (kind[k] == 1) {
              mnx = min(mnx, f2ord(rx0[k])); mxx = max(mxx, f2ord(rx1[k]));
              mny = min(mny, f2ord(ry0[k])); mxy = max(mxy, f2ord(ry1[k]));
            }
            if constexpr (kZ) {
              rz0[k] = rz1[k] = 0.f;
              zkind[k] = kind[k] == 0 ? 0 : box_zslab(bq[k][2], bq[k][5], rz0[k], rz1[k]);
              if (zkind[k] == 1) { mnz = min(mnz, f2ord(rz0[k])); mxz = max(mxz, f2ord(rz1[k])); }
            }
          } else if constexpr (kZ) {
            zkind[k] = 0;
            rz0[k] = rz1[k] = 0.f;
          }
        }
        if constexpr (kZ) { mnz = __reduce_min_sync(0xffffffffu, mnz); mxz = __reduce_max_sync(0xffffffffu, mxz); }
        mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
        mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
        if (lane == 0) {
          s_wext[warp][0] = mnx; s_wext[warp][1] = mny; s_wext[warp][2] = mxx; s_wext[warp][3] = mxy;
          if constexpr (kZ) { s_wext[warp][4] = mnz; s_wext[warp][5] = mxz; }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * nbw) : "memory");  // warp extents published
        request_rest();
        stamp();
        // bin function of this sweep: the inner bins tile the bounding rectangle of the finite box
        // rectangles; any non-negative scale is valid (points and boxes share it)
        uint32_t e0 = 0xffffffffu, e1 = 0xffffffffu, e2 = 0u, e3 = 0u;
        if (lane < nbw) { e0 = s_wext[lane][0]; e1 = s_wext[lane][1]; e2 = s_wext[lane][2]; e3 = s_wext[lane][3]; }
        e0 = __reduce_min_sync(0xffffffffu, e0); e1 = __reduce_min_sync(0xffffffffu, e1);
        e2 = __reduce_max_sync(0xffffffffu, e2); e3 = __reduce_max_sync(0xffffffffu, e3);
        if (e0 != 0xffffffffu) {
          const float gx0 = ord2f(e0), gy0 = ord2f(e1);
          const float wx = ord2f(e2) - gx0, wy = ord2f(e3) - gy0;
          if (wx > 0.f && isfinite(wx)) invx = __fdividef((float)(kBins - 2), wx);
          if (wy > 0.f && isfinite(wy)) invy = __fdividef((float)(kBins - 2), wy);
          if (!isfinite(invx) || !(invx >= 0.f)) invx = 0.f;
          if (!isfinite(invy) || !(invy >= 0.f)) invy = 0.f;
          offx = __fsub_rn(1.f, __fmul_rn(gx0, invx));  // bin = floor(v * inv + off): bin 1 starts at g0
          offy = __fsub_rn(1.f, __fmul_rn(gy0, invy));
          if (!isfinite(offx)) { invx = 0.f; offx = 1.f; }
          if (!isfinite(offy)) { invy = 0.f; offy = 1.f; }
        }
        if constexpr (kZ) {  // the same for the z slabs: kZBins - 2 inner bins tile their extent
          uint32_t z0 = 0xffffffffu, z1 = 0u;
          if (lane < nbw) { z0 = s_wext[lane][4]; z1 = s_wext[lane][5]; }
          z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
          if (z0 != 0xffffffffu) {
            const float gz0 = ord2f(z0), wz = ord2f(z1) - gz0;
            if (wz > 0.f && isfinite(wz)) invz = __fdividef((float)(kZBins - 2), wz);
            if (!isfinite(invz) || !(invz >= 0.f)) invz = 0.f;
            offz = __fsub_rn(1.f, __fmul_rn(gz0, invz));
            if (!isfinite(offz)) { invz = 0.f; offz = 1.f; }
          }
        }
        if (tid == 0) {
          s_bin[0] = invx; s_bin[1] = offx; s_bin[2] = invy; s_bin[3] = offy;
          if constexpr (kZ) { s_bin[4] = invz; s_bin[5] = offz; }
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if constexpr (kZ) {
            bz0[k] = 0; bz1[k] = kZBins - 1;                      // zkind 2: every z bin
            if (zkind[k] == 1) { bz0[k] = zbin_of(rz0[k], invz, offz); bz1[k] = zbin_of(rz1[k], invz, offz); }
            if (zkind[k] == 0) { bz0[k] = 1; bz1[k] = 0; }        // none (but the last bin, below)
          }
          bx0[k] = 0; bx1[k] = kBins - 1; by0[k] = 0; by1[k] = kBins - 1;
          if (kind[k] == 1) {
            bx0[k] = bin_of(rx0[k], invx, offx); bx1[k] = bin_of(rx1[k], invx, offx);
            by0[k] = bin_of(ry0[k], invy, offy); by1[k] = bin_of(ry1[k], invy, offy);
          }
          if constexpr (kCells) {
            const int t = tid + k * NT;
            if (t < Tc)  // an empty rectangle is stored as x0 > x1
              s_rect[t] = kind[k] == 0 ? 1u : ((uint32_t)bx0[k] | ((uint32_t)bx1[k] << 8) | ((uint32_t)by0[k] << 16) | ((uint32_t)by1[k] << 24));
          }
        }
      }
      if (!box_warp) {
        if (terms_warp) {  // the boxes this warp needs have arrived once their first value is usable
          if (__float_as_uint(bq[0][6]) == 0x7fc12345u) s_bin[0] = 0.f;  // (never true: orders the request after the load)
        }
        request_rest();
        if (terms_warp && eoff != 0) exact_terms();  // concurrently with the box warps (they fit beside them)
        stamp();
      }
      bool need_masks = true;
      if constexpr (kCells) {
        // sparse path: every box appends its id to the cells (2 x 2 bins) its rectangle touches; the
        // (box, cell row) pairs are spread over ALL threads: 2^lp threads per box, strided rows
        __syncthreads();  // rectangles published (and the exact terms done)
        stamp();
        int lp = 0;
        while (lp < 3 && (Tc << (lp + 1)) <= NT) ++lp;
        const int P = 1 << lp;
        for (int u = tid; u < (Tc << lp); u += NT) {
          const uint32_t t = (uint32_t)(u >> lp);
          const int part = u & (P - 1);
          const uint32_t rc = s_rect[t];
          if ((rc & 0xffu) > ((rc >> 8) & 0xffu)) continue;
          const int cx0 = cell_of_bin((int)(rc & 0xffu)), cx1 = cell_of_bin((int)((rc >> 8) & 0xffu));
          const int cy0 = cell_of_bin((int)((rc >> 16) & 0xffu)), cy1 = cell_of_bin((int)(rc >> 24));
          if ((cx1 - cx0 + 1) * (cy1 - cy0 + 1) > kWideCells) {  // a few huge boxes: a short list tested for every point
            if (part == 0) {
              const int sl = atomicAdd(&s_nwide, 1);
              if (sl < kMaxWide) s_wide[sl] = (int)t;
              else s_flags[1] = 1;
            }
            continue;
          }
          for (int cy = cy0 + part; cy <= cy1; cy += P) {
            uint32_t* rowp = cells + 2 * (cy * kCellsSide);
            for (int cx = cx0; cx <= cx1; cx += 4) {
              uint32_t sl[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)  // the count lives in the top byte of word 1; four cells in flight
                sl[j] = cx + j <= cx1 ? atomicAdd(rowp + 2 * (cx + j) + 1, 1u << 24) >> 24 : 0u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (cx + j > cx1) continue;
                uint32_t* e = rowp + 2 * (cx + j);
                if (sl[j] < 4u) atomicOr(e, t << (8u * sl[j]));
                else if (sl[j] < (uint32_t)kCellIds) atomicOr(e + 1, t << (8u * (sl[j] - 4u)));
                else s_flags[0] = 1;
                if (sl[j] >= 254u) s_flags[1] = 1;  // the one-byte count is about to wrap: do not trust the cells
              }
            }
          }
        }
        stamp();
        __syncthreads();
        need_masks = (s_flags[0] | s_flags[1]) != 0;  // dense spots fall back to the axis masks
      }
      if (box_warp) {
        if (need_masks) {
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            if (kind[k] == 0) continue;
            const int t = tid + k * NT;
            const uint32_t bit = 1u << (t & 31);
            uint32_t* cx_ = mx + (t >> 5);
            uint32_t* cy_ = my + (t >> 5);
            for (int b = bx0[k]; b <= bx1[k]; ++b) atomicOr(cx_ + b * ms, bit);
            for (int b = by0[k]; b <= by1[k]; ++b) atomicOr(cy_ + b * ms, bit);
            if constexpr (kZ) {
              uint32_t* cz_ = mz + (t >> 5);
              for (int b = bz0[k]; b <= bz1[k]; ++b) atomicOr(cz_ + b * ms, bit);
              // the last bin takes every point above the slabs and every NaN z — and a NaN z passes
              // the z test of EVERY box (the CPU op's `fabsf(z - cz) > dz/2` is false for NaN)
              if (bz1[k] != kZBins - 1) atomicOr(cz_ + (kZBins - 1) * ms, bit);
            }
          }
        }
        stamp();
        if (eoff == 0) exact_terms();
      }
      if (!kCells || need_masks || eoff == 0) __syncthreads();  // (all three conditions are CTA-uniform)
      invx = s_bin[0]; offx = s_bin[1]; invy = s_bin[2]; offy = s_bin[3];
      if constexpr (kZ) { invz = s_bin[4]; offz = s_bin[5]; }
#ifdef GGA_PROFILING
      if ((p.variant & 512) && tj == 0 && lane == 0)
        for (int k = 1; k < S; ++k) request_tile(k, tr + R * k);
#endif
      stamp();

      // ---- sweep ---------------------------------------------------------------------------
      const int nck = wc >> 2;  // 16-byte chunks per row (wide rows)
      // Lane l writes the chunks of its row in the rotated order (j + rot) mod nck: with rows of 32,
      // 64 or 128 bytes the 8 lanes of a quarter warp then hit 8 different bank groups although the
      // stage is the plain linear image of the rows (which the bulk store needs).
      const int rot = nck > 0 ? ((lane * nck) >> 3) % nck : 0;
      bool bulk_ok = kBulk && p.W == wc && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0;
#ifdef GGA_PROFILING
      if (p.variant & 32) bulk_ok = false;
#endif
      int buf = 0;
      const bool cells_bad = kCells && s_flags[1] != 0;
      const int n_wide = kCells ? min(s_nwide, kMaxWide) : 0;
      auto batch = [&](const float4 pt, const int gb) {
        const int bx = bin_of(pt.x, invx, offx), by = bin_of(pt.y, invy, offy);
        uint32_t* stage = stage_base + (kStages == 2 ? buf * 32 * wcap : 0);
        if constexpr (kBulk) {
          if (bulk_ok) {  // the engine must be done READING this buffer
            if (lane == 0) bulk_wait_read<kStages - 1>();
            __syncwarp();
          }
        }
        uint32_t* row = stage + lane * wc;
        int best = -1;
        if constexpr (kCells) {
          // ---- sparse path: one 8-byte cell entry names the candidates; the row starts empty and the
          //      boxes that pass the exact test set their bit
          const uint2 e = *reinterpret_cast<const uint2*>(cells + 2 * (cell_of_bin(by) * kCellsSide + cell_of_bin(bx)));
          if constexpr (kWide) {
#pragma unroll
            for (int j = 0; j < WC / 4; ++j) {
              int k = j + rot;
              if (k >= nck) k -= nck;
              reinterpret_cast<uint4*>(row)[k] = make_uint4(0u, 0u, 0u, 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < WC; ++k) row[k] = 0u;
          }
          auto test_box = [&](const uint32_t t) {
            const float4 a = terms[2 * t];
            if (outside_z(pt.z, a)) return;
            if (!inside_xy(pt.x, pt.y, a, terms[2 * t + 1])) return;
            if constexpr (MODE == kModePart) best = best < 0 ? (int)t : min(best, (int)t);
            else atomicOr(row + (t >> 5), 1u << (t & 31u));  // lane-private row: the atomic is just the shortest RMW
          };
          uint32_t cnt = e.y >> 24;
#ifdef GGA_PROFILING
          if (p.variant & 1) cnt = 0u;
#endif
          if (cnt > (uint32_t)kCellIds || cells_bad) {
            // dense spot: every box the axis masks name
#pragma unroll 1
            for (int w = 0; w < wc; ++w) {
              uint32_t m = mx[bx * ms + w] & my[by * ms + w];
              while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1u;
                test_box((uint32_t)((w << 5) + j));
              }
            }
          } else {
            unsigned long long ids = ((unsigned long long)(e.y & 0x00ffffffu) << 32) | e.x;
#pragma unroll 1
            for (; cnt != 0u; --cnt) {
              test_box((uint32_t)(ids & 0xffull));
              ids >>= 8;
            }
#pragma unroll 1
            for (int i = 0; i < n_wide; ++i) test_box((uint32_t)s_wide[i]);
          }
        } else {
          // ---- general path: candidate row = maskx[bx] & masky[by], written to the lane's stage row
          //      (nz = its non-zero words); the exact test clears the bits that fail
          uint32_t nz = 0u;
          const uint4* ax = reinterpret_cast<const uint4*>(mx + bx * ms);
          const uint4* ay = reinterpret_cast<const uint4*>(my + by * ms);
          const uint4* az = reinterpret_cast<const uint4*>(mz + (kZ ? zbin_of(pt.z, invz, offz) : 0) * ms);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j >= nck) break;
            int k = j + rot;
            if (k >= nck) k -= nck;
            uint4 a = ax[k];
            const uint4 b = ay[k];
            a.x &= b.x; a.y &= b.y; a.z &= b.z; a.w &= b.w;
            if constexpr (kZ) {
              const uint4 c = az[k];
              a.x &= c.x; a.y &= c.y; a.z &= c.z; a.w &= c.w;
            }
            reinterpret_cast<uint4*>(row)[k] = a;
            nz |= ((a.x != 0u ? 1u : 0u) | (a.y != 0u ? 2u : 0u) | (a.z != 0u ? 4u : 0u) | (a.w != 0u ? 8u : 0u)) << (4 * k);
          }
#ifdef GGA_PROFILING
          if (p.variant & 1) nz = 0u;
#endif
          // one candidate per loop trip (lanes diverge only in the trip count)
          if (nz != 0u) {
            int w = __ffs(nz) - 1;
            uint32_t m = row[w], keep = m;
            for (;;) {
              const int j = __ffs(m) - 1;
              m &= m - 1u;
              const int t = (w << 5) + j;
              const float4 a = terms[2 * t];
              bool ok = !outside_z(pt.z, a);
              if (ok) ok = inside_xy(pt.x, pt.y, a, terms[2 * t + 1]);
              if constexpr (MODE == kModePart) {
                if (ok) { best = t0 + t; break; }
              } else {
                if (!ok) keep &= ~(1u << j);
              }
              if (m == 0u) {
                if constexpr (MODE != kModePart) row[w] = keep;
                nz &= nz - 1u;
                if (nz == 0u) break;
                w = __ffs(nz) - 1;
                m = keep = row[w];
              }
            }
          }
        }

        const int nvalid = min(32, N - gb * 32);
        const size_t row0 = (size_t)f * N + (size_t)gb * 32;  // first point of the batch
        if constexpr (MODE == kModePart) {
          if (lane < nvalid) {
            int32_t* dst = reinterpret_cast<int32_t*>(p.out) + row0 + lane;
            if (ch == 0) __stcs(dst, best);
            else if (best >= 0 && *dst < 0) *dst = best;  // later sweeps: only where no earlier box matched
          }
        } else if constexpr (MODE == kModeBits && !kWide) {
          if (lane < nvalid) {  // one row per lane: the warp store is contiguous already
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.out) + (row0 + lane) * WC;
            if constexpr (WC == 1) __stcs(dst, row[0]);
            else __stcs(reinterpret_cast<uint2*>(dst), *reinterpret_cast<const uint2*>(row));
          }
        } else if constexpr (MODE == kModeBits) {
          uint32_t* base = reinterpret_cast<uint32_t*>(p.out) + row0 * p.W + 32 * ch;
#ifdef GGA_PROFILING
          if (p.variant & 8) {
            __syncwarp();
          } else
#endif
          if (bulk_ok) {  // rows are contiguous in memory: the batch is one linear run of bytes
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store(base, stage, (uint32_t)(nvalid * wc * 4));
            buf ^= 1;
          } else {
            __syncwarp();
            const int total = nvalid * nck;  // 16-byte chunks of the batch
            if (p.W == wc) {
              uint4* dst = reinterpret_cast<uint4*>(base);
#pragma unroll 1
              for (int c = lane; c < total; c += 32) __stcs(dst + c, reinterpret_cast<const uint4*>(stage)[c]);
            } else {  // more than 1024 boxes: a sweep writes a 128-byte slice of every row
#pragma unroll 1
              for (int c = lane; c < total; c += 32) {
                const int q = c / nck, k = c - q * nck;
                __stcs(reinterpret_cast<uint4*>(base + (size_t)q * p.W) + k, reinterpret_cast<const uint4*>(stage)[c]);
              }
            }
            __syncwarp();
          }
        } else {  // kModeAll: int32 [N, T] rows; lane l writes boxes 4l .. 4l+3 (+128 k) of every point
          __syncwarp();
          const int T = p.T;
          int32_t* dst = reinterpret_cast<int32_t*>(p.out) + row0 * T + t0;
          if ((T & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) {
#pragma unroll 1
            for (int q = 0; q < nvalid; ++q)
              for (int t4 = lane * 4; t4 < Tc; t4 += 128) {
                const uint32_t nib = stage[q * wc + (t4 >> 5)] >> (t4 & 31);
                __stcs(reinterpret_cast<int4*>(dst + (size_t)q * T + t4),
                       make_int4(nib & 1u, (nib >> 1) & 1u, (nib >> 2) & 1u, (nib >> 3) & 1u));
              }
          } else {
#pragma unroll 1
            for (int q = 0; q < nvalid; ++q)
              for (int t = lane; t < Tc; t += 32)
                __stcs(dst + (size_t)q * T + t, (int32_t)((stage[q * wc + (t >> 5)] >> (t & 31)) & 1u));
          }
          __syncwarp();
        }
      };

      // this warp's tiles: frame tile ft0 + k * dft, ring slot tr + R * (k mod S); everything advanced incrementally
      const int dft = rf * R;
      int ft = rr + rf * tr;
      int ks = 0;        // ring slot of this warp group's current tile: tr + R * ks
      uint32_t ph = 0u;  // mbarrier phase of the current pass over the ring
      const uint32_t full0 = smem_u32(&s_full[tr]), empty0 = smem_u32(&s_empty[tr]);
      const unsigned char* tile0 = ring + (size_t)tr * tile_bytes;
      const int pin = tj * 32 + lane;  // this lane's point inside a tile
#pragma unroll 1
      for (int k = 0; ft < ntiles; ++k, ft += dft) {
        const int gb = ft * kTileBatches + tj;
        const bool ringed = ft < n_full;
        float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ringed) {
          mbar_wait_a(full0 + (uint32_t)(ks * R) * 8u, ph);
          const unsigned char* tile = tile0 + (size_t)(ks * R) * tile_bytes;
          if constexpr (VEC4) {
            pt = reinterpret_cast<const float4*>(tile)[pin];
          } else {
            const float* q = reinterpret_cast<const float*>(tile) + (size_t)pin * p.pts_stride;
            pt = make_float4(q[0], q[1], q[2], 0.f);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_a(empty0 + (uint32_t)(ks * R) * 8u);
        } else if (gb < bpf) {
          pt = load_pt(gb);
        }
        if (gb < bpf) batch(pt, gb);
        if (ringed && tj == 0 && lane == 0 && ft + dft * S < n_full) {
          mbar_wait_a(empty0 + (uint32_t)(ks * R) * 8u, ph);  // the 8 warps of this tile have read their points
          request_tile(k + S, tr + R * ks);
        }
        if (++ks == S) { ks = 0; ph ^= 1u; }
      }
      if constexpr (kBulk) {
        if (lane == 0) bulk_wait_all();  // the stage buffers are rewritten (or released) next
        __syncwarp();
      }
      stamp();
    }
  }
}

// Compacts bit-packed rows into a list of (point, box) pairs: one warp per 32 rows, the pairs of a
// warp are appended with one atomicAdd (the order of the list is unspecified; the count is exact even
// when the list overflows its capacity, pairs beyond it are dropped).
__global__ void __launch_bounds__(256) hit_list_kernel(const uint32_t* __restrict__ bits, long long rows, int W,
                                                       long long row_base, int32_t* __restrict__ pairs, int capacity,
                                                       int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  for (long long r0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32; r0 < rows; r0 += (long long)gridDim.x * 256) {
    const long long r = r0 + lane;
    int n = 0;
    if (r < rows)
      for (int w = 0; w < W; ++w) n += __popc(__ldg(bits + r * W + w));
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    int base = 0;
    if (lane == 31) base = atomicAdd(count, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    if (n) {
      for (int w = 0; w < W; ++w) {
        uint32_t m = __ldg(bits + r * W + w);
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1u;
          if (base < capacity) {
            pairs[2 * (long long)base] = (int32_t)(row_base + r);
            pairs[2 * (long long)base + 1] = (w << 5) + j;
          }
          ++base;
        }
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

#ifdef GGA_PROFILING
struct ProfKnobs {
  int nt = 0;      // 0 = default, else 512 / 1024 threads per CTA
  int ranges = 0;  // 0 = default, else CTAs per frame
  int variant = 0;
  unsigned long long* trace = nullptr;
};
ProfKnobs g_prof;
#endif

// shared memory of a CTA without the point ring
size_t sweep_fixed_smem(int mode, int T, int W, int nt) {
  const size_t tcap = T < kChunkBoxes ? T : kChunkBoxes;
  const size_t wcap = W < 32 ? W : 32;
  const size_t stages = (mode == kModeBits && W >= 4 && W <= 8) ? 2 : 1;  // double-buffered narrow rows (bulk store)
  const size_t cells = W <= 8 ? (size_t)kCellsSide * kCellsSide * 8 : 0;  // sparse path (<= 256 boxes)
  const size_t mask_rows = 2 * kBins + (W > 8 && nt == 1024 ? kZBins : 0);  // general path at 1024 threads: z mask too
  return tcap * 32 + mask_rows * mask_stride((int)wcap) * 4 + (size_t)(nt / 32) * stages * 32 * wcap * 4 + cells;
}

template <int MODE, int WC, int NT, bool VEC4>
int launch_sweep(const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = pib_sweep_kernel<MODE, WC, NT, VEC4>;
  if (smem > 48 * 1024) {
    // idempotent and cheap; no cached "already configured" state to keep the library re-entrant
    GGA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GGA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return GGA_OK;
}

template <int MODE, bool VEC4>
int launch_by_width(const SweepParams& p, int nt, dim3 grid, size_t smem, cudaStream_t st) {
  if (nt == 1024) {
    switch (p.W) {
      case 1: return launch_sweep<MODE, 1, 1024, VEC4>(p, grid, smem, st);
      case 2: return launch_sweep<MODE, 2, 1024, VEC4>(p, grid, smem, st);
      case 4: return launch_sweep<MODE, 4, 1024, VEC4>(p, grid, smem, st);
      case 8: return launch_sweep<MODE, 8, 1024, VEC4>(p, grid, smem, st);
      default: return launch_sweep<MODE, 0, 1024, VEC4>(p, grid, smem, st);
    }
  }
#ifdef GGA_PROFILING
  if (p.W == 8) return launch_sweep<MODE, 8, 512, VEC4>(p, grid, smem, st);
#endif
  return launch_sweep<MODE, 0, 512, VEC4>(p, grid, smem, st);  // any multiple of 8 words
}

template <int MODE>
int launch_mode(const SweepParams& p, int nt, bool vec4, dim3 grid, size_t smem, cudaStream_t st) {
  return vec4 ? launch_by_width<MODE, true>(p, nt, grid, smem, st) : launch_by_width<MODE, false>(p, nt, grid, smem, st);
}

int run_pib(int mode, const float* points, int pts_stride, const float* boxes, void* out, int B, int num_points,
            int num_boxes, void* stream) {
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
  GGA_REQUIRE(num_points <= (1 << 30), "at most 2^30 points per frame");

  SweepParams sp;
  sp.points = points; sp.boxes = boxes; sp.out = out;
  sp.pts_stride = pts_stride; sp.N = num_points; sp.T = num_boxes; sp.B = B;
  sp.W = gga_pib_row_words(num_boxes);
  sp.bpf = (num_points + 31) / 32;
  sp.nchunks = (num_boxes + kChunkBoxes - 1) / kChunkBoxes;
  sp.trace = nullptr;
  sp.variant = 0;
  // 1024-thread CTAs (one per SM, 32 warps) for rows up to 8 words, and for bit rows of 16 / 24 words
  // (257..768 boxes: the sweep is bound by dependent latency per batch, twice the warps hide more of
  // it — measured 70 -> 56 us at 8 x 50k x 512); 512 threads for 32-word rows and the int32 layouts,
  // whose per-warp stage leaves too little room for the point ring next to 32 warps (slower there)
  int nt = sp.W <= 8 || (mode == kModeBits && sp.W <= 24) ? 1024 : 512;
#ifdef GGA_PROFILING
  if (g_prof.nt == 512 && sp.W >= 8) nt = 512;
  if (g_prof.nt == 1024) nt = 1024;
  sp.trace = g_prof.trace;
  sp.variant = g_prof.variant;
#endif
  const size_t fixed = sweep_fixed_smem(mode, num_boxes, sp.W, nt);
  const size_t limit = (size_t)gga_max_smem_optin() - 4096;  // minus the kernel's static shared memory (~2.6 KB) and slack
  GGA_REQUIRE(fixed <= limit, "internal: %zu bytes of shared memory requested", fixed);
  const int nsm = gga_sm_count();
  // CTAs per frame: fill the machine (one CTA per SM) but keep at least one batch per warp; more
  // frames than SMs: one CTA per frame, the CTAs loop over the frames
  const int nwarps = nt / 32;
  long long rf = B >= nsm ? 1 : nsm / B;
  const long long rf_max = (sp.bpf + nwarps - 1) / nwarps;
  if (rf > rf_max) rf = rf_max;
  if (rf < 1) rf = 1;
#ifdef GGA_PROFILING
  if (g_prof.ranges > 0) rf = g_prof.ranges < sp.bpf ? g_prof.ranges : sp.bpf;
#endif
  int gy = B;
  if ((long long)gy * rf > nsm) gy = (int)(nsm / rf) > 0 ? (int)(nsm / rf) : 1;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)rf, (unsigned)gy);
  // point ring: as many 256-point tiles as a CTA will consume, bounded by the shared memory left;
  // a multiple of the tiles consumed at a time (one per 8 warps)
  {
    const int R = nwarps / kTileBatches;
    const size_t tile_bytes = (size_t)1024 * pts_stride;
    const long long ntiles = (sp.bpf + kTileBatches - 1) / kTileBatches;
    long long want = (ntiles + rf - 1) / rf;  // tiles of one frame per CTA
    want = (want + R - 1) / R * R;
    long long fit = (long long)((limit - fixed) / tile_bytes) / R * R;
    if (fit > kMaxSlots / R * R) fit = kMaxSlots / R * R;
    sp.ring_slots = (int)(want < fit ? want : fit);
    if (pts_stride > 16 || sp.ring_slots < R) sp.ring_slots = 0;
  }
  const size_t smem = fixed + (size_t)sp.ring_slots * 1024 * pts_stride;
  const bool vec4 = pts_stride == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0;
  if (mode == kModeBits) return launch_mode<kModeBits>(sp, nt, vec4, grid, smem, st);
  if (mode == kModeAll) return launch_mode<kModeAll>(sp, nt, vec4, grid, smem, st);
  return launch_mode<kModePart>(sp, nt, vec4, grid, smem, st);
}

}  // namespace

extern "C" int gga_pib_row_words(int num_boxes) {
  if (num_boxes <= 0) return 0;
  if (num_boxes <= 32) return 1;
  if (num_boxes <= 64) return 2;
  if (num_boxes <= 128) return 4;
  return 8 * ((num_boxes + 255) / 256);
}

extern "C" int gga_points_in_boxes_bits(const float* points, int pts_stride, const float* boxes, uint32_t* bits,
                                        int B, int num_points, int num_boxes, void* stream) {
  return run_pib(kModeBits, points, pts_stride, boxes, bits, B, num_points, num_boxes, stream);
}

extern "C" int gga_points_in_boxes_all(const float* points, int pts_stride, const float* boxes, int32_t* out,
                                       int B, int num_points, int num_boxes, void* stream) {
  return run_pib(kModeAll, points, pts_stride, boxes, out, B, num_points, num_boxes, stream);
}

extern "C" int gga_points_in_boxes_part(const float* points, int pts_stride, const float* boxes, int32_t* out,
                                        int B, int num_points, int num_boxes, void* stream) {
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

extern "C" int gga_pib_hit_list(const uint32_t* bits, int64_t num_rows, int num_boxes, int64_t row_base,
                                int32_t* pairs, int capacity, int32_t* count, int reset_count, void* stream) {
  GGA_REQUIRE(num_rows >= 0 && num_boxes >= 0 && capacity >= 0, "negative size");
  GGA_REQUIRE(count != nullptr, "null count pointer");
  cudaStream_t st = gga_stream(stream);
  if (reset_count) GGA_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
  if (num_rows == 0 || num_boxes == 0) return GGA_OK;
  GGA_REQUIRE(bits != nullptr && (pairs != nullptr || capacity == 0), "null pointer");
  const int W = gga_pib_row_words(num_boxes);
  long long blocks = (num_rows + 255) / 256;
  const long long cap_blocks = (long long)gga_sm_count() * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  hit_list_kernel<<<(unsigned)blocks, 256, 0, st>>>(bits, num_rows, W, row_base, pairs, capacity, count);
  GGA_CHECK_CUDA(cudaGetLastError());
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

#ifdef GGA_PROFILING
/* developer builds only (tools/build_prof.py): CTA size / ranges per frame overrides and a
 * per-CTA phase timeline.  Not part of the product library. */
extern "C" int gga_prof_pib(int nt, int ranges_per_frame, int variant, void* trace_buffer) {
  g_prof.nt = nt;
  g_prof.ranges = ranges_per_frame;
  g_prof.variant = variant;
  g_prof.trace = static_cast<unsigned long long*>(trace_buffer);
  return GGA_OK;
}
#endif
