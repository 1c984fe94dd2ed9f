// SLAB feed of the implicit-GEMM convolution (tcgen05.mma, fp32 accumulators in TMEM).
//
// An output tile is 16 rows x (8*tm) columns of one image (tm = 1 or 2 "halves" of 128 GEMM rows
// each: row m = 8*g + r  <->  output pixel (y0 + g, x0 + 8*half + r)).  For every K stage the
// producer issues ONE 5-D TMA box that drops the input slab the tile needs (tile + filter reach)
// into shared memory; every filter tap is then just a UMMA shared-memory descriptor whose start
// address is shifted inside that slab:
//
//   TEDSPAD_SLAB_3X3     slab rows = pixels x 64 channels (128 B, SWIZZLE_128B); an 8-row core group
//                        = 8 consecutive x of one image row; SBO = slab row pitch; tap (ky,kx) =
//                        start + (ky*slab_w + kx) * 128 B.  Input bytes cross L2->SM once per tile
//                        instead of once per tap (9x less operand traffic than the FLAT feed).
//   TEDSPAD_SLAB_STEM*   un-swizzled K-major descriptors whose K-adjacent core matrices OVERLAP
//                        (LBO = 16 B): GEMM row m, K chunk j reads the 16 bytes at (m + j) * 16 B of
//                        a slab row, i.e. the im2col window of a small-Cin convolution in place.
//
// The weights live in shared memory for the whole life of the persistent CTA, already in the layout
// the B descriptors read (tedspad_conv_slab_pack); the MMA issuer walks a host-built table of
// (A offset, B offset) pairs, so the kernel itself knows nothing about filter geometry.
//
//   warps 0-7   epilogue       two warps per TMEM lane quarter; tcgen05.ld (software-pipelined 32-column
//                              chunks) -> +bias -> ReLU -> bf16 stores; optional fused MaxPool2d(2)
//                              (warp shuffles) and OutConv 1x1 + sigmoid
//   warp 8      TMA producer   weights once (cp.async.bulk), then one slab box per (tile, K stage)
//   warp 9      MMA issuer     warp-uniform loop, one elected lane issues n_mma x tm tcgen05.mma per K stage
//   warp 10     TMEM allocator
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"
#include "ptx.cuh"

namespace tsp {

constexpr int SLAB_SMEM_BUDGET = 227 * 1024;
constexpr int SLAB_TAIL_BYTES = 2048 /*bias*/ + 1024 /*outconv w,b*/ + 512 /*barriers*/;
constexpr int SLAB_MAX_BSTAGES = 8;
constexpr int SLAB_MAX_STAGES = 6;
constexpr int SLAB_MAX_ACC = 4;      // TMEM accumulator ring depth (2..4: as many as fit in the 512 columns)

// n / d for any 32-bit n >= 0 with magic = ceil(2^64 / d) (0 encodes d == 1): exact because n * d < 2^64.  A hardware
// integer division is ~25 instructions; the epilogue warps decode (n_tile, tx, ty, tz, image, row) for every tile,
// and those warps run at one instruction per ~5 clocks (ncu: 20 % issue-selected).
struct DivMagic {
  uint64_t m;
  int d;
};
__device__ __forceinline__ int fdiv(int n, const DivMagic& k) {
  return k.m ? static_cast<int>(__umul64hi(static_cast<uint64_t>(static_cast<uint32_t>(n)), k.m)) : n;
}
static DivMagic make_div(int d) {
  DivMagic k;
  k.d = d;
  k.m = d <= 1 ? 0 : (~0ULL / static_cast<uint64_t>(d)) + 1;   // = ceil(2^64 / d) for every d >= 2
  return k;
}

struct SlabKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;           // streamed weights: standard packed [Cout_pad][K_pad], box {64, n_tile}
  const uint8_t* w_image;
  const float* bias;
  int tm, n_tile, k_stages, n_grp, nk, a_kstep, b_kstep, stages, tmem_cols, acc_stages, bias_floats;
  int slab_bytes, slab_stride, w_bytes, w_stride, zero_slabs;
  int b_stream, b_stages, b_stride, cb_n, cin, num_n_tiles, tab_per_stage;
  // fused x2 bilinear up-sampling source (channel blocks >= up_cb_first are interpolated into the slab)
  const __nv_bfloat16* up;
  int up_cb_first, up_H, up_W, up_Hp, up_Wp, up_ph, up_pw, up_ld, up_coff, up_UH, up_UW, up_offy, up_offx, slab_w, slab_px;
  float up_sy, up_sx;
  uint32_t slab_w_magic;
  int half_a_off;
  int c_step, x_step, x_off, y_step, y_off, z_step, z_off, z_kstep, merged_cw;
  int tiles_x, tiles_y, tiles_z, total_tiles;
  // tile index -> (n_tile, a, b, tz, image) with a fastest; (a, b) = (tx, ty), or (ty, tx) when ty_first: CTA pairs
  // then work on tiles of the SAME column block, so that a last column block whose second 8-column half lies outside
  // the image (W = 56: 7 groups) can skip that half's MMAs for both CTAs of the pair at once
  int ty_first, dim_a, dim_b;
  // KX kind: accumulator columns = 3 filter columns x kx_cp output channels (slab_epilogue_kx); bias_n = bias floats
  int kx, kx_cp, bias_n;
  DivMagic dv_nt, dv_tx, dv_ty, dv_tz, dv_hp;   // divisors: num_n_tiles, dim_a, dim_b, tiles_z, stack_hp
  int stack_hp, stack_ph, stack_n;   // stacked rows (see make_plan): padded image height, halo rows, batch; 0 = off
  uint64_t a_desc, b_desc;
  // epilogue
  __nv_bfloat16* y;
  int OH, OW, yDp, yHp, yWp, ypd, yph, ypw, y_ld, y_coff, Cout, act;
  __nv_bfloat16* pool;
  int PH, PW, pHp, pWp, pph, ppw, p_ld, p_coff;
  const float* oc_w;
  const float* oc_b;
  __nv_bfloat16* oc_planes;
  float* oc_frames;
  const __nv_bfloat16* res;    // optional bf16 residual with y's geometry (added before the activation)
  int res_ld, res_coff, res_wide;   // res_wide: every 32-channel chunk of a pixel is 32-byte aligned (256-bit loads)
  int res_prefetch;                 // L2 prefetch of the next tile's residual rows (epi_res_prefetch_tile)
  // staged stores (EPI_PLAIN, tm == 2, n_tile == 64, single CTA): every epilogue warp writes its 32 pixels x 128 B into a
  // swizzled shared-memory tile and one lane issues a TMA tensor store of full 128-byte lines (profiles/
  // r2j_hbm_write_probe.md: one-row-per-lane st.global tops out at 4.8 TB/s, TMA bulk stores reach 6.3)
  CUtensorMap tmY;
  int tst, tst_off;                 // on / byte offset of the 8 warps x 2 x 4 KB staging area in shared memory
  __nv_bfloat16* oc_clip;      // encoder clip written through the raw-reshape glue (NULL = planes only)
  int oc_T, cDp, cHp, cWp, cpd, cph, cpw, c_ld, c_coff;
  DivMagic dv_T;
  uint2 tab[TEDSPAD_SLAB_MAX_MMA];
};

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t max_bf162(uint32_t a, uint32_t b) {
  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}

// r / d for r < 2^16 with magic = ceil(2^32 / d)
__device__ __forceinline__ int fast_div16(int r, uint32_t magic) { return static_cast<int>(__umulhi(static_cast<uint32_t>(r), magic)); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// MMA issue, executed by ALL lanes of the issuing warp with warp-uniform values (so the compiler keeps the
// descriptors in uniform registers); one elected lane issues the tcgen05 instructions.  A K stage is n_grp
// table groups (one filter tap / filter row each) of NK K=16 steps; the next group's table entry is fetched
// while the current group's MMAs are issued.  TM and NK are compile-time so that the inner loop is straight
// line code of ~3 uniform instructions per tcgen05.mma: at N = 64 the tensor pipe wants a new instruction
// every ~48 clocks and a single warp retires a dependent uniform instruction only every ~7.
template <int TM, int NK, bool STREAM, bool PAIR>
__device__ __forceinline__ void slab_issue(const SlabKParams& p, uint8_t* smS, uint32_t w_addr, uint32_t tmem_base,
                                           uint64_t* full, uint64_t* empty, uint64_t* tfull, uint64_t* tempty,
                                           uint64_t* wbar, uint64_t* bfull, uint64_t* bempty) {
  const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.n_tile);   // PAIR: M = 256 over the two CTAs
  const uint32_t a_hi = static_cast<uint32_t>(p.a_desc >> 32), a_lo0 = static_cast<uint32_t>(p.a_desc);
  const uint32_t b_hi = static_cast<uint32_t>(p.b_desc >> 32);
  const uint32_t b_lo0 = static_cast<uint32_t>(p.b_desc) + (w_addr >> 4);
  // swizzled kinds (every 3x3 / 1x1 kind: 128-byte pixel rows): K=16 step = 32 B, second half = 8 pixels further.  As
  // compile-time constants they fold into the descriptor adds (the issuer of an N = 64 layer has ~32 clocks per MMA)
  constexpr bool SWZ = STREAM || NK == 4;
  const uint32_t half_step = SWZ ? 64u : static_cast<uint32_t>(p.half_a_off >> 4);
  const uint32_t a_ks = SWZ ? 2u : static_cast<uint32_t>(p.a_kstep >> 4), b_ks = SWZ ? 2u : static_cast<uint32_t>(p.b_kstep >> 4);
  const int S = p.stages, NG = p.n_grp, KS = p.k_stages, total = p.total_tiles;
  const uint32_t n_tile = static_cast<uint32_t>(p.n_tile);
  const uint32_t slab_step = static_cast<uint32_t>(p.slab_stride >> 4);
  const uint32_t a_lo_base = a_lo0 + (smem_u32(smS) >> 4);
  const uint32_t b_step = static_cast<uint32_t>(p.b_stride >> 4);
  const int BS = p.b_stages, tab_ps = p.tab_per_stage;
  if (!STREAM) mbar_wait(wbar, 0);
  if (PAIR && !STREAM) mbar_wait_cluster(bfull, 0);   // the peer's half of the weights has landed in ITS shared memory
  int s = 0, as = 0, bs = 0;
  uint32_t ph = 0, aph = 0, bph = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    mbar_wait(tempty + as, aph ^ 1);   // PAIR: 16 arrivals, the peer's come through mbar_arrive_cluster
    tc_fence_after();
    const uint32_t d0 = tmem_base + static_cast<uint32_t>(as * TM) * n_tile;
    const uint32_t d1 = d0 + n_tile;
    // the second 8-column half of the last tile of a row may lie entirely outside the image (W = 56: 7 groups):
    // its MMAs are skipped (its epilogue warps find nothing valid to store)
    bool h1 = TM == 2;
    if (TM == 2 && !PAIR) h1 = ((tile / p.num_n_tiles) % p.tiles_x) * 16 + 8 < p.OW;
    if (TM == 2 && PAIR && p.ty_first) {
      // one M = 256 instruction covers half 1 of BOTH tiles of the pair (tile, tile + 1): skip it when neither has one
      const int t0 = tile / p.num_n_tiles, t1 = (tile + 1) / p.num_n_tiles;
      const int tx0 = (t0 / p.dim_a) % p.dim_b, tx1 = (t1 / p.dim_a) % p.dim_b;
      h1 = tx0 * 16 + 8 < p.OW || tx1 * 16 + 8 < p.OW;
    }
    for (int ks = 0; ks < KS; ++ks) {
      const uint2* tab = p.tab + (tab_ps ? ks * NG : 0);
      uint2 cur = tab[0];
      mbar_wait(full + s, ph);
      tc_fence_after();
      const uint32_t a_lo_s = a_lo_base + static_cast<uint32_t>(s) * slab_step;
      for (int g = 0; g < NG; ++g) {
        const uint2 nxt = tab[g + 1 < NG ? g + 1 : g];
        const uint32_t a_lo = a_lo_s + cur.x;
        uint32_t b_lo;
        if (STREAM) {  // this tap's weight block arrives through its own ring
          mbar_wait(bfull + bs, bph);
          tc_fence_after();
          b_lo = b_lo0 + static_cast<uint32_t>(bs) * b_step;
        } else {
          b_lo = b_lo0 + cur.y;
        }
        const uint32_t acc0 = (g | ks) != 0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * a_ks);
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * b_ks);
            if (PAIR) umma_bf16_nc_pair(d0, ad, bd, idesc, k == 0 ? acc0 : 1u);
            else umma_bf16_nc(d0, ad, bd, idesc, k == 0 ? acc0 : 1u);
            if (TM == 2 && h1) {
              const uint64_t ad1 = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * a_ks + half_step);
              if (PAIR) umma_bf16_nc_pair(d1, ad1, bd, idesc, k == 0 ? acc0 : 1u);
              else umma_bf16_nc(d1, ad1, bd, idesc, k == 0 ? acc0 : 1u);
            }
          }
          if (STREAM) { if (PAIR) umma_commit_pair(bempty + bs); else umma_commit(bempty + bs); }
        }
        if (STREAM) {
          __syncwarp();
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
        cur = nxt;
      }
      if (elect_one()) { if (PAIR) umma_commit_pair(empty + s); else umma_commit(empty + s); }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one()) { if (PAIR) umma_commit_pair(tfull + as); else umma_commit(tfull + as); }
    __syncwarp();
    if (++as == p.acc_stages) { as = 0; aph ^= 1; }
  }
}

// One 32-column chunk of one half of the tile, for this thread's pixel: +bias, activation, bf16 stores and the
// optional fused epilogues.  v = raw fp32 accumulators from TMEM.
struct EpiCtx {
  const float* bias;   // shared
  const float* ocw;    // shared: [3][Cout] then [3]
  __nv_bfloat16* y;
  __nv_bfloat16* pool;
  int Cout, act, y_ld, y_coff, p_ld, p_coff;
  bool fuse_oc, wide_ok, res_wide;
  const __nv_bfloat16* res;   // EPI_RES: bf16 residual with y's pixel geometry
  int res_ld, res_coff;
  uint8_t* stg;               // EPI_PLAIN staged stores: this warp's 4 KB tile (row = lane), or nullptr
  int stg_c0;                 // first channel of the N tile (staging column 0)
};

// MODE: 0 = bf16 stores only, 1 = + fused MaxPool2d(2), 2 = fused OutConv (stores / pool optional at run time),
// 3 = fused OutConv alone (the last UNet layer: its 64-channel tensor is never written).
// Compile-time modes keep run-time branches out of the per-chunk code: the epilogue warps are latency-bound (ncu:
// issue-selected 12-16 % of their samples, the rest short-scoreboard / fixed-latency waits).  (Processing both
// chunks of a 64-output tile as one instruction stream - no TMEM-load / math overlap, twice the live registers -
// was measured slower: 64->64 + OutConv 1.70 -> 2.20 ms.)
enum { EPI_PLAIN = 0, EPI_POOL = 1, EPI_OC = 2, EPI_OC_ONLY = 3, EPI_RES = 4 };   // OC_ONLY: OutConv with no 64-channel
                                                                                  // store, no pool; RES: plain + residual

// The 64 bytes (32 bf16) of residual of one chunk of this thread's pixel, fetched with pinned loads when the chunk's
// accumulator load is issued (so that the DRAM round trip overlaps the previous chunk's arithmetic).
__device__ __forceinline__ void epi_res_fetch(const EpiCtx& c, long long pix, int c0, bool valid, uint4 (&r)[4]) {
  const __nv_bfloat16* rp = c.res + pix * c.res_ld + c.res_coff + c0;
  if (c.res_wide && c0 + 32 <= c.Cout) {
    // two 32-byte loads: a warp load instruction touches 32 cache lines whatever its width (one pixel per lane), so the
    // residual costs half the L1 wavefronts of four 16-byte loads - the epilogue of the K = 64 bottleneck tails is
    // bound by exactly those (1x1 64 -> 256 + residual: 2.3x the time of the same layer without)
    r[0] = r[1] = r[2] = r[3] = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      ld_nc_256_pinned(rp, r[0], r[1]);
      ld_nc_256_pinned(rp + 16, r[2], r[3]);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r[j] = make_uint4(0u, 0u, 0u, 0u);
    if (valid && c0 + 8 * j < c.Cout) r[j] = ld_nc_v4_pinned(rp + 8 * j);
  }
}

template <int MODE, bool RELU>
__device__ __forceinline__ void epi_math(const EpiCtx& c, const uint32_t (&v)[32], int c0, uint32_t (&q)[16], float (&oc)[3],
                                         const uint4 (&r)[4]) {
  if (MODE != EPI_OC && MODE != EPI_OC_ONLY) {
    // 16 packed fp32x2 bias adds + 16 converts with the ReLU fused into the rounding instruction
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(&r[0]);   // EPI_RES: 16 packed bf16 pairs
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(c.bias + c0 + i);
      float f0 = __uint_as_float(v[i]), f1 = __uint_as_float(v[i + 1]);
      float f2 = __uint_as_float(v[i + 2]), f3 = __uint_as_float(v[i + 3]);
      add_f32x2(f0, f1, b4.x, b4.y);
      add_f32x2(f2, f3, b4.z, b4.w);
      if (MODE == EPI_RES) {   // bn3 + residual, then ReLU (large_i3d.py:72-79); a bf16 is the upper half of an fp32
        const uint32_t p0 = rw[i >> 1], p1 = rw[(i >> 1) + 1];
        add_f32x2(f0, f1, __uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u));
        add_f32x2(f2, f3, __uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
      }
      q[i >> 1] = cvt_bf16x2(f0, f1, RELU);
      q[(i >> 1) + 1] = cvt_bf16x2(f2, f3, RELU);
    }
  } else {
    // fused OutConv 1x1: the dot products consume the un-rounded fp32 activations
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(c.bias + c0 + i);
      f[i] = __uint_as_float(v[i]); f[i + 1] = __uint_as_float(v[i + 1]);
      f[i + 2] = __uint_as_float(v[i + 2]); f[i + 3] = __uint_as_float(v[i + 3]);
      add_f32x2(f[i], f[i + 1], b4.x, b4.y);
      add_f32x2(f[i + 2], f[i + 3], b4.z, b4.w);
    }
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (c0 < c.Cout) {
      // packed fp32x2 FMAs into two independent accumulator pairs per output
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float* wr = c.ocw + o * c.Cout + c0;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wr + i);
          fma_f32x2(a0, a1, f[i], f[i + 1], w4.x, w4.y);
          fma_f32x2(b0, b1, f[i + 2], f[i + 3], w4.z, w4.w);
        }
        oc[o] += (a0 + a1) + (b0 + b1);
      }
    }
    if (MODE == EPI_OC && (c.y != nullptr || c.pool != nullptr)) {
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] = pack_bf162(f[2 * i], f[2 * i + 1]);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void epi_out(const EpiCtx& c, const uint32_t (&q)[16], int c0, bool valid, long long pix,
                                        bool pool_writer, long long ppix) {
  if (MODE == EPI_OC_ONLY) return;
  if (MODE == EPI_PLAIN && c.stg != nullptr) {
    // row = lane (128 bytes = 64 channels), 16-byte chunk k at position k ^ (row & 7): the SWIZZLE_128B image the
    // tensor map of the store expects, and conflict-free for the 8 lanes of a quarter warp
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t* row = c.stg + lane * 128u;
    const uint32_t k0 = static_cast<uint32_t>(c0 - c.stg_c0) >> 3;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(row + (((k0 + j) ^ (lane & 7u)) << 4)) = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
    return;
  }
  if ((MODE != EPI_OC || c.y != nullptr) && valid) {
    // each lane owns 64 contiguous bytes of its pixel: two 32-byte stores (a warp store instruction touches 32
    // cache lines whatever its width, so wider stores halve the L1 wavefronts per byte)
    __nv_bfloat16* yp = c.y + pix * c.y_ld + c.y_coff + c0;
    if (c.wide_ok && c0 + 32 <= c.Cout) {
      st_global_256(yp, q);
      st_global_256(yp + 16, q + 8);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + 8 * j < c.Cout)
          *reinterpret_cast<uint4*>(yp + 8 * j) = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
    }
  }
  if (MODE == EPI_POOL || (MODE == EPI_OC && c.pool != nullptr)) {
    // MaxPool2d(2): x partner = lane ^ 1 (r ^ 1), y partner = lane ^ 8 (g ^ 1).  Each exchange also halves the
    // channels a lane carries on (the partner keeps the other half), so the 2x2 window costs 8 + 4 shuffles per
    // 32 channels instead of 32, and all four lanes of the window store 16 bytes (8 channels) of the pooled pixel.
    const uint32_t lane = threadIdx.x & 31u;
    const bool odd_x = (lane & 1u) != 0, odd_y = (lane & 8u) != 0;
    // (all shuffles of an exchange are issued before the first max: written as shuffle-then-max pairs, every max
    // waited out its own shuffle - ncu: 25 % of the epilogue's samples on the 24 HMNMX2)
    uint32_t a[8], b[4], t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = __shfl_xor_sync(0xffffffffu, odd_x ? q[i] : q[i + 8], 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = max_bf162(odd_x ? q[i + 8] : q[i], t[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = __shfl_xor_sync(0xffffffffu, odd_y ? a[i] : a[i + 4], 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = max_bf162(odd_y ? a[i + 4] : a[i], t[i]);
    const int cq = c0 + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);   // first of this lane's 8 pooled channels
    if (pool_writer && cq < c.Cout)
      *reinterpret_cast<uint4*>(c.pool + ppix * c.p_ld + c.p_coff + cq) = make_uint4(b[0], b[1], b[2], b[3]);
  }
}

// The epilogue role (warps 0-7).  Two warps per TMEM lane quarter.  tm == 2: warp group eg owns half eg of the tile
// (all its columns, so the fused OutConv dot product stays inside one thread); tm == 1: the groups split the
// 32-column chunks.
//
// E16 (64-output CTA-pair layers: one K stage of MMAs per tile is SHORTER than the 8-warp epilogue of that tile, ncu
// profiles/r2_ncu_64ch_layers.md): sixteen warps, four per TMEM lane quarter = four per scheduler, each owning ONE
// 32-column chunk of one half (warp >> 2 = 2 * half + chunk).  A warp hands its accumulator back as soon as its single
// tcgen05.ld has landed in registers, i.e. before any of the arithmetic; the fused OutConv dot product of a pixel is
// completed by the chunk-0 warp from the chunk-1 warp's partial sums (shared-memory scratch, one named barrier per
// warp pair and tile, double-buffered by tile parity).
// EPI_RES: L2 prefetch of the residual rows of `tile` (this lane's pixel, the tile's N-tile channel range), issued one
// tile ahead of their use.  The bottleneck tails (1x1, K = 64..512) have an epilogue that is never idle, so the
// two-chunk register prefetch of epi_res_fetch covers only a fraction of a DRAM round trip; an L2 prefetch needs no
// registers and turns those loads into L2 hits.  Tile decode as in slab_epilogue.
__device__ __forceinline__ void epi_res_prefetch_tile(const SlabKParams& p, int tile, int g, int r, int h, int eg, int tm) {
  int t = tile, q;
  q = fdiv(t, p.dv_nt);
  const int n0 = (t - q * p.dv_nt.d) * p.n_tile; t = q;
  q = fdiv(t, p.dv_tx);
  const int ta = t - q * p.dv_tx.d; t = q;
  q = fdiv(t, p.dv_ty);
  const int tb = t - q * p.dv_ty.d; t = q;
  const int tx = p.ty_first ? tb : ta, ty = p.ty_first ? ta : tb;
  q = fdiv(t, p.dv_tz);
  const int tz = t - q * p.dv_tz.d;
  int n = q;
  int oy = ty * 16 + g;
  const int ox = (tx * tm + h) * 8 + r;
  bool valid = ox < p.OW;
  if (p.stack_hp) {
    const int R = oy + p.stack_ph;
    n = fdiv(R, p.dv_hp);
    oy = R - n * p.stack_hp - p.stack_ph;
    valid = valid && n < p.stack_n;
  }
  valid = valid && static_cast<unsigned>(oy) < static_cast<unsigned>(p.OH);
  if (!valid) return;
  const long long pix = ((static_cast<long long>(n) * p.yDp + tz + p.ypd) * p.yHp + oy + p.yph) * p.yWp + ox + p.ypw;
  const __nv_bfloat16* rp = p.res + pix * p.res_ld + p.res_coff + n0;
  const int nc = min(p.n_tile, p.Cout - n0);                  // real channels of this N tile
  // tm == 1: the two warp groups share the pixels (alternate 32-channel chunks): each takes every second line
  const int first = tm == 1 ? 64 * eg : 0, step = tm == 1 ? 128 : 64;
  for (int ch = first; ch < nc; ch += step) prefetch_l2(rp + ch);
}

template <bool HAS_UP, bool PAIR, int MODE, bool RELU, bool E16>
__device__ __forceinline__ void slab_epilogue(const SlabKParams& p, int warp, int lane, uint32_t tmem_base, const float* sm_bias,
                                              const float* sm_ocw, uint64_t* tfull, uint64_t* tempty, float* sm_part,
                                              uint8_t* sm_stage) {
  pdl_wait();   // the previous kernel may still be reading the buffers this one writes
  const int ew = warp & 3, eg = E16 ? (warp >> 3) : (warp >> 2);
  const int g = ew * 4 + (lane >> 3);
  const int r = lane & 7;
  const int tm = p.tm, n_tile = p.n_tile, OH = p.OH, OW = p.OW;
  const int nchunk_all = n_tile >> 5;
  EpiCtx c;
  c.bias = sm_bias; c.ocw = sm_ocw; c.y = p.y; c.pool = p.pool; c.Cout = p.Cout; c.act = p.act;
  c.y_ld = p.y_ld; c.y_coff = p.y_coff; c.p_ld = p.p_ld; c.p_coff = p.p_coff; c.fuse_oc = (MODE == EPI_OC || MODE == EPI_OC_ONLY);
  // 32-byte stores need 32-byte aligned pixel chunks
  c.wide_ok = ((p.y_ld | p.y_coff) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
  c.res = p.res; c.res_ld = p.res_ld; c.res_coff = p.res_coff;
  c.res_wide = p.res_wide != 0;
  c.stg = nullptr; c.stg_c0 = 0;
  const bool tst = MODE == EPI_PLAIN && !HAS_UP && !PAIR && !E16 && p.tst != 0;   // (host: tm == 2, n_tile == 64, no stacking)
  int tst_buf = 0;
  int h, c_first, c_step, nch;
  if (E16) { h = eg; c_first = 32 * ((warp >> 2) & 1); c_step = 64; nch = 1; }   // tm == 2, n_tile == 64
  else if (tm == 2) { h = eg; c_first = 0; c_step = 32; nch = nchunk_all; }
  else if ((MODE == EPI_OC || MODE == EPI_OC_ONLY)) { h = 0; c_first = 0; c_step = 32; nch = eg == 0 ? nchunk_all : 0; }
  else { h = 0; c_first = 32 * eg; c_step = 64; nch = (nchunk_all + 1 - eg) >> 1; }
  int as = 0, tpar = 0;
  uint32_t aph = 0;
  const bool res_pf = MODE == EPI_RES && !E16 && p.res_prefetch != 0;
  if (res_pf && static_cast<int>(blockIdx.x) < p.total_tiles) epi_res_prefetch_tile(p, blockIdx.x, g, r, h, eg, tm);
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    if (res_pf && tile + static_cast<int>(gridDim.x) < p.total_tiles)
      epi_res_prefetch_tile(p, tile + gridDim.x, g, r, h, eg, tm);   // one tile ahead
    int t = tile, q;
    q = fdiv(t, p.dv_nt);
    const int n0 = (t - q * p.dv_nt.d) * n_tile; t = q;   // first output channel of this N tile
    q = fdiv(t, p.dv_tx);
    const int ta = t - q * p.dv_tx.d; t = q;
    q = fdiv(t, p.dv_ty);
    const int tb = t - q * p.dv_ty.d; t = q;
    const int tx = p.ty_first ? tb : ta, ty = p.ty_first ? ta : tb;
    q = fdiv(t, p.dv_tz);
    const int tz = t - q * p.dv_tz.d;
    int n = q;
    int oy = ty * 16 + g;
    const int ox = (tx * tm + h) * 8 + r;
    bool valid = ox < OW;
    if (p.stack_hp) {
      // stacked rows: the tile's 16 rows are consecutive rows of the zero-haloed images of the whole batch
      const int R = oy + p.stack_ph;
      n = fdiv(R, p.dv_hp);
      oy = R - n * p.stack_hp - p.stack_ph;
      valid = valid && n < p.stack_n;
    }
    valid = valid && static_cast<unsigned>(oy) < static_cast<unsigned>(OH);
    const long long pix = ((static_cast<long long>(n) * p.yDp + tz + p.ypd) * p.yHp + oy + p.yph) * p.yWp + ox + p.ypw;
    const int py = oy >> 1, px = ox >> 1;
    // all four lanes of a 2x2 window write (8 channels each); the window is in the image iff its pooled pixel is
    const bool pool_writer = oy >= 0 && ox < OW && py < p.PH && px < p.PW && (p.stack_hp == 0 || n < p.stack_n);
    const long long ppix = (static_cast<long long>(n) * p.pHp + py + p.pph) * p.pWp + px + p.ppw;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                           static_cast<uint32_t>((as * tm + h) * n_tile + c_first);
    float oc[3] = {0.f, 0.f, 0.f};
    if (tst) {
      // the store that read this staging buffer two tiles ago must have finished reading it
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      c.stg = sm_stage + (warp * 2 + tst_buf) * 4096;
      c.stg_c0 = n0;
    }
    mbar_wait(tfull + as, aph);
    tc_fence_after();
    uint4 ra[4], rb[4];   // EPI_RES only (dead otherwise)
    if (E16) {
      uint32_t va[32], qa[16];
      const int c0 = n0 + c_first;
      tmem_ld32(t_row, va);
      if (MODE == EPI_RES) epi_res_fetch(c, pix, c0, valid, ra);
      tmem_ld_wait();
      // this warp's only read of the accumulator is complete: hand it back before the arithmetic
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(tempty + as), 0));
        else mbar_arrive(tempty + as);
      }
      epi_math<MODE, RELU>(c, va, c0, qa, oc, ra);
      epi_out<MODE>(c, qa, c0, valid, pix, pool_writer, ppix);
      if (MODE == EPI_OC || MODE == EPI_OC_ONLY) {
        // partial 64 -> 3 dot products of the upper chunk travel to the lower chunk's warp (same pixels)
        // (both warps block on the pair's barrier, so the writer is never more than one tile ahead of the reader:
        // two scratch buffers alternating by tile parity are enough)
        float* part = sm_part + ((tpar * 8 + h * 4 + ew) * 32 + lane) * 3;
        const int bar_id = 1 + h * 4 + ew;
        if (c_first != 0) { part[0] = oc[0]; part[1] = oc[1]; part[2] = oc[2]; }
        __syncwarp();   // bar.sync is .aligned: the lanes diverged on `valid` above must have reconverged
        named_bar_sync(bar_id, 64);
        if (c_first == 0) { oc[0] += part[0]; oc[1] += part[1]; oc[2] += part[2]; }
        tpar ^= 1;
      }
    } else if (HAS_UP) {   // 480-thread variant: 136 registers per thread, one chunk in flight
      uint32_t va[32], qa[16];
      for (int i = 0; i < nch; ++i) {
        const int c0 = n0 + c_first + i * c_step;
        tmem_ld32(t_row + i * c_step, va);
        if (MODE == EPI_RES) epi_res_fetch(c, pix, c0, valid, ra);
        tmem_ld_wait();
        epi_math<MODE, RELU>(c, va, c0, qa, oc, ra);
        epi_out<MODE>(c, qa, c0, valid, pix, pool_writer, ppix);
      }
    } else {
      // software pipeline over this warp's chunks: the next chunk's TMEM load (and residual) is in flight while one
      // is processed
      uint32_t va[32], vb[32], qa[16];
      const int cbase = n0 + c_first;
      if (nch > 0) {
        tmem_ld32(t_row, va);
        if (MODE == EPI_RES) epi_res_fetch(c, pix, cbase, valid, ra);
      }
      for (int i = 0; i < nch; i += 2) {
        tmem_ld_wait();
        const int c0 = cbase + i * c_step;
        if (i + 1 < nch) {
          tmem_ld32(t_row + (i + 1) * c_step, vb);
          if (MODE == EPI_RES) epi_res_fetch(c, pix, c0 + c_step, valid, rb);
        }
        epi_math<MODE, RELU>(c, va, c0, qa, oc, ra);
        epi_out<MODE>(c, qa, c0, valid, pix, pool_writer, ppix);
        if (i + 1 < nch) {
          tmem_ld_wait();
          if (i + 2 < nch) {
            tmem_ld32(t_row + (i + 2) * c_step, va);
            if (MODE == EPI_RES) epi_res_fetch(c, pix, c0 + 2 * c_step, valid, ra);
          }
          epi_math<MODE, RELU>(c, vb, c0 + c_step, qa, oc, rb);
          epi_out<MODE>(c, qa, c0 + c_step, valid, pix, pool_writer, ppix);
        }
      }
    }
    // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
    if (!E16) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(tempty + as), 0));
        else mbar_arrive(tempty + as);
      }
    }
    if (tst) {
      // the warp's 4 rows x 8 columns x 64 channels leave as ONE tensor store (box {64, 8, 4, 1, 1}); pixels outside the
      // image are outside the tensor map and are not written
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tma_store_5d(&p.tmY, c.stg, n0, (tx * tm + h) * 8, ty * 16 + ew * 4, tz, n);
      tst_buf ^= 1;
    }
    if ((MODE == EPI_OC || MODE == EPI_OC_ONLY) && valid && nch > 0 && (!E16 || c_first == 0)) {
      const long long plane = static_cast<long long>(OH) * OW;
      const long long o0 = static_cast<long long>(n) * 3 * plane + static_cast<long long>(oy) * OW + ox;
      const int bclip = fdiv(n, p.dv_T), tf = n - bclip * p.oc_T;   // frame n = clip bclip, time tf
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float sgm = 1.f / (1.f + __expf(-(oc[o] + sm_ocw[3 * p.Cout + o])));
        const __nv_bfloat16 sb = __float2bfloat16_rn(sgm);
        if (p.oc_planes != nullptr) p.oc_planes[o0 + o * plane] = sb;
        if (p.oc_frames != nullptr) p.oc_frames[o0 + o * plane] = sgm;
        if (p.oc_clip != nullptr) {
          // dali_extraction.py:171-173: plane 3*tf + o of the clip's 3T planes is encoder channel ce at time te
          const int pl = 3 * tf + o, ce = fdiv(pl, p.dv_T), te = pl - ce * p.oc_T;
          const long long cp = ((static_cast<long long>(bclip) * p.cDp + te + p.cpd) * p.cHp + oy + p.cph) * p.cWp + ox + p.cpw;
          p.oc_clip[cp * p.c_ld + p.c_coff + ce] = sb;
        }
      }
    }
    if (++as == p.acc_stages) { as = 0; aph ^= 1; }
  }
  if (tst && lane == 0) bulk_wait_all<0>();   // the staging area must outlive the stores that read it
}

// KX kinds (TEDSPAD_SLAB_3X3_KX_PAIR).  At N <= 64 outputs the MMA is bound by its A-operand reads: a 3x3 layer reads
// every slab pixel nine times from shared memory, once per tap.  Here the three taps of one filter ROW share a fetch:
// B stacks the weights of kx = 0, 1, 2 along N (N = 3 * CP), so one K step of filter row ky computes, for slab pixel
// q = (row, s), the three partial sums  D_kx[q] = sum_c w[ky][kx][c] * x[row + ky - 1][s][c]  - 3 fetches per pixel
// instead of 9 - and the output pixel (row, j) is  D_0[(row, j)] + D_1[(row, j + 1)] + D_2[(row, j + 2)]:  a tile is
// 8 rows x 16 slab columns (128 GEMM rows, flat: the slab is exactly 16 pixels wide, so an 8-pixel core group is half
// a slab row and SBO = 1024 B) and yields 14 output columns; the two neighbour terms come from the next lanes of
// the warp (a warp holds two slab rows of 16 lanes each).  Warp >> 2 selects the half of the CP channels.
template <bool PAIR, bool RELU>
__device__ __forceinline__ void slab_epilogue_kx(const SlabKParams& p, int warp, int lane, uint32_t tmem_base, const float* sm_bias,
                                                 uint64_t* tfull, uint64_t* tempty) {
  pdl_wait();   // the previous kernel may still be reading the buffers this one writes
  const int ew = warp & 3, eg = warp >> 2;
  const int CP = p.kx_cp, CH = CP >> 1, c_lo = eg * CH;   // channels per filter column block / per warp; first channel
  const int row = 2 * ew + (lane >> 4), s = lane & 15;
  const bool wide_ok = ((p.y_ld | p.y_coff) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
  int as = 0;
  uint32_t aph = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int t = tile, q;
    q = fdiv(t, p.dv_tx);
    const int tx = t - q * p.dv_tx.d; t = q;
    q = fdiv(t, p.dv_ty);
    const int ty = t - q * p.dv_ty.d;
    int n = q;                      // 2-D layers: tiles_z == 1
    int oy = ty * 8 + row;
    const int ox = tx * 14 + s;
    bool valid = s < 14 && ox < p.OW;
    if (p.stack_hp) {
      const int R = oy + p.stack_ph;
      n = fdiv(R, p.dv_hp);
      oy = R - n * p.stack_hp - p.stack_ph;
      valid = valid && n < p.stack_n;
    }
    valid = valid && static_cast<unsigned>(oy) < static_cast<unsigned>(p.OH);
    const long long pix = ((static_cast<long long>(n) * p.yDp + p.ypd) * p.yHp + oy + p.yph) * p.yWp + ox + p.ypw;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as * p.n_tile + c_lo);
    mbar_wait(tfull + as, aph);
    tc_fence_after();
    for (int c = 0; c < CH; c += 16) {
      uint32_t d0[16], d1[16], d2[16];
      tmem_ld16(t_row + c, d0);
      tmem_ld16(t_row + CP + c, d1);
      tmem_ld16(t_row + 2 * CP + c, d2);
      tmem_ld_wait();
      if (c + 16 >= CH) {   // last read of this accumulator: hand it back before the arithmetic
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(tempty + as), 0));
          else mbar_arrive(tempty + as);
        }
      }
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float f0 = __uint_as_float(d0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(d1[i]), 1, 16) +
                   __shfl_down_sync(0xffffffffu, __uint_as_float(d2[i]), 2, 16);
        float f1 = __uint_as_float(d0[i + 1]) + __shfl_down_sync(0xffffffffu, __uint_as_float(d1[i + 1]), 1, 16) +
                   __shfl_down_sync(0xffffffffu, __uint_as_float(d2[i + 1]), 2, 16);
        const float2 b2 = *reinterpret_cast<const float2*>(sm_bias + c_lo + c + i);
        f0 += b2.x; f1 += b2.y;
        o[i >> 1] = cvt_bf16x2(f0, f1, RELU);
      }
      const int cg = c_lo + c;   // first output channel of this chunk
      if (p.oc_clip != nullptr && valid && cg == 0) {
        // space-to-depth head of the UNet++ anonymizer -> encoder clip through the raw-reshape glue (dali_extraction.py:
        // 171-173), what frames_to_clip_kernel(s2d) does from memory: channel (2a + b) * 3 + colour of low-res pixel
        // (oy, ox) of frame n is colour `colour` of pixel (2 oy + a, 2 ox + b); plane 3 tf + colour of the clip's 3T
        // planes is encoder channel ce at time te
        const int bclip = fdiv(n, p.dv_T), tf = n - bclip * p.oc_T;
        long long cbase[3];
#pragma unroll
        for (int col = 0; col < 3; ++col) {
          const int pl = 3 * tf + col, ce = fdiv(pl, p.dv_T), te = pl - ce * p.oc_T;
          cbase[col] = (((static_cast<long long>(bclip) * p.cDp + te + p.cpd) * p.cHp + 2 * oy + p.cph) * p.cWp + 2 * ox + p.cpw) *
                           p.c_ld + p.c_coff + ce;
        }
        const long long row_step = static_cast<long long>(p.cWp) * p.c_ld;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          const int ab = k / 3, col = k - 3 * ab;
          const uint16_t v = static_cast<uint16_t>((k & 1) ? (o[k >> 1] >> 16) : (o[k >> 1] & 0xffffu));
          reinterpret_cast<uint16_t*>(p.oc_clip)[cbase[col] + (ab >> 1) * row_step + (ab & 1) * p.c_ld] = v;
        }
      }
      if (p.y != nullptr && valid && cg < p.Cout) {
        __nv_bfloat16* yp = p.y + pix * p.y_ld + p.y_coff + cg;
        if (wide_ok && cg + 16 <= p.Cout) {
          st_global_256(yp, o);
        } else {
          *reinterpret_cast<uint4*>(yp) = make_uint4(o[0], o[1], o[2], o[3]);
          if (cg + 8 < p.Cout) *reinterpret_cast<uint4*>(yp + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    if (++as == p.acc_stages) { as = 0; aph ^= 1; }
  }
}

// The four bilinear taps (16 bytes = 8 channels each) and weights of slab pixel r of the up-sampled half.
struct UpTaps {
  uint4 qa, qb, qc, qd;
  float lx0, lx1, ly0, ly1;
  bool in;
};

__device__ __forceinline__ UpTaps up_fetch(const SlabKParams& p, int r, int iy0, int ix0, int n, int ch) {
  UpTaps t;
  const int sy = fast_div16(r, p.slab_w_magic), sx = r - sy * p.slab_w;
  const int iy = iy0 + sy, ix = ix0 + sx;
  const int uy = iy - p.up_offy, ux = ix - p.up_offx;
  t.in = static_cast<unsigned>(iy) < static_cast<unsigned>(p.OH) && static_cast<unsigned>(ix) < static_cast<unsigned>(p.OW) &&
         static_cast<unsigned>(uy) < static_cast<unsigned>(p.up_UH) && static_cast<unsigned>(ux) < static_cast<unsigned>(p.up_UW);
  if (t.in) {
    const float fy = p.up_sy * uy, fx = p.up_sx * ux;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < p.up_H - 1 ? 1 : 0), x1 = x0 + (x0 < p.up_W - 1 ? 1 : 0);
    t.ly1 = fy - y0; t.ly0 = 1.f - t.ly1; t.lx1 = fx - x0; t.lx0 = 1.f - t.lx1;
    const long long rb0 = (static_cast<long long>(n) * p.up_Hp + y0 + p.up_ph) * p.up_Wp + p.up_pw;
    const long long rb1 = (static_cast<long long>(n) * p.up_Hp + y1 + p.up_ph) * p.up_Wp + p.up_pw;
    t.qa = __ldg(reinterpret_cast<const uint4*>(p.up + (rb0 + x0) * p.up_ld + ch));
    t.qb = __ldg(reinterpret_cast<const uint4*>(p.up + (rb0 + x1) * p.up_ld + ch));
    t.qc = __ldg(reinterpret_cast<const uint4*>(p.up + (rb1 + x0) * p.up_ld + ch));
    t.qd = __ldg(reinterpret_cast<const uint4*>(p.up + (rb1 + x1) * p.up_ld + ch));
  }
  return t;
}

constexpr int SLAB_THREADS = 352;      // warps 0-7 epilogue (two per TMEM lane quarter), 8 producer, 9 MMA, 10 TMEM alloc
constexpr int SLAB_THREADS_UP = 480;   // + warps 11-14: up-sampling slab producers
constexpr int SLAB_THREADS_E16 = 608;  // warps 0-15 epilogue (four per TMEM lane quarter), 16 producer, 17 MMA, 18 TMEM alloc
constexpr int SLAB_E16_SCRATCH = 2 * 8 * 32 * 3 * 4;   // OutConv partial sums: [tile parity][half][quarter][lane][3] floats

// HAS_UP: the input is the concatenation [skip | upsample2x(low-res)] of Up.forward (unet_parts.py:57-67) and the
// up-sampled half is never materialised: for its channel blocks four producer warps interpolate the low-res
// tensor (bilinear, align_corners=True, F.pad offsets) straight into the swizzled slab the TMA would have
// filled.  Same arithmetic as upsample2x_kernel (ops.cu), so fused == unfused bit for bit.
//
// PAIR: the grid is made of clusters of two CTAs (the two SMs of a TPC) that run one tcgen05.mma.cta_group::2 of
// M = 256 per step: CTA r of the pair owns tile blockIdx.x (its own slab, its own accumulator, its own epilogue)
// and HALF of the output-channel rows of the weights.  At N = 64 a single-CTA MMA is bound by shared-memory
// reads ((128 + 64) rows x 32 B per 32 tensor clocks = 67 % of the tensor peak, profiles/r1_umma_probe.txt);
// with the B rows split over two SMs it is (128 + 32) rows = 80 %.  Only the leader issues MMAs; its `full`
// barriers collect the TMA bytes of both CTAs, commits are multicast to both, the peer's epilogue warps
// arrive remotely on the leader's `tempty`.
template <bool HAS_UP, bool PAIR, bool E16>
__global__ void __launch_bounds__(HAS_UP ? SLAB_THREADS_UP : (E16 ? SLAB_THREADS_E16 : SLAB_THREADS), 1)
conv_slab_kernel(const __grid_constant__ SlabKParams p) {
  constexpr int NEW = E16 ? 16 : 8;                      // epilogue warps
  constexpr int W_PROD = NEW, W_MMA = NEW + 1, W_ALLOC = NEW + 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // round up inside the shared window (pointer arithmetic on the symbol keeps the shared address space,
  // so bias / OutConv weights are read with LDS instead of generic loads)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int S = p.stages;
  uint8_t* smW = smem;
  uint8_t* smS = smem + p.w_stride;
  float* sm_bias = reinterpret_cast<float*>(smS + S * p.slab_stride);
  float* sm_ocw = sm_bias + p.bias_floats;   // [3][Cout] then [3] bias
  uint64_t* full = reinterpret_cast<uint64_t*>(sm_ocw + 256);
  uint64_t* empty = full + SLAB_MAX_STAGES;
  uint64_t* tfull = empty + SLAB_MAX_STAGES;
  uint64_t* tempty = tfull + SLAB_MAX_ACC;
  uint64_t* wbar = tempty + SLAB_MAX_ACC;
  uint64_t* bfull = wbar + 1;
  uint64_t* bempty = bfull + SLAB_MAX_BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + SLAB_MAX_BSTAGES);
  float* sm_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 512);   // E16 only (make_plan reserves it)

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // == blockIdx.x & 1: the tile loops below need no change

  if (warp == W_PROD && lane == 0) {
    if (p.tst) tma_prefetch_desc(&p.tmY);
    tma_prefetch_desc(&p.tmA);
    if (p.b_stream) tma_prefetch_desc(&p.tmB);
  }
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, HAS_UP ? 1 + 4 : 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < p.acc_stages; ++a) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, PAIR ? 2 * NEW : NEW);   // PAIR (leader): the epilogue warps of both CTAs
    }
    mbar_init(wbar, 1);
    if (PAIR && !p.b_stream) mbar_init(bfull, 1);   // leader: "the peer's weights are resident"
    for (int b = 0; b < p.b_stages; ++b) {
      mbar_init(bfull + b, 1);
      mbar_init(bempty + b, 1);
    }
    fence_barrier_init();
  }
  if (warp == W_ALLOC) {
    if (PAIR) {
      // one warp of EACH CTA of the pair executes the collective cta_group::2 allocation with the SAME shared-memory
      // offset (per-rank slots were tried to silence compute-sanitizer racecheck, which reports the pair's two
      // identical-value writes as a hazard: the hardware then leaves one CTA's slot unwritten -> misaligned-address
      // trap in its epilogue.  The racecheck report on this line is a tool artefact, see profiles/r2_sanitizer.md)
      tmem_alloc_pair(tmem_slot, p.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, p.tmem_cols);
      tmem_relinquish();
    }
  }
  for (int i = threadIdx.x; i < p.bias_n; i += blockDim.x) sm_bias[i] = p.bias[i];
  if (p.oc_w != nullptr) {
    for (int i = threadIdx.x; i < 3 * p.Cout; i += blockDim.x) sm_ocw[i] = p.oc_w[i];
    if (threadIdx.x < 3) sm_ocw[3 * p.Cout + threadIdx.x] = p.oc_b[threadIdx.x];
  }
  if (p.zero_slabs) {
    // un-swizzled stems read a few bytes past the TMA box (zero-weight K padding): keep them finite
    uint4* z = reinterpret_cast<uint4*>(smS);
    const int n16 = S * p.slab_stride / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its prologue on SMs this grid has left (common.h)

  if (warp == W_PROD) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (!p.b_stream) {
        const uint8_t* w_src = p.w_image + (PAIR ? static_cast<size_t>(rank) * p.w_bytes : 0);   // this CTA's half
        mbar_arrive_expect_tx(wbar, static_cast<uint32_t>(p.w_bytes));
        for (int off = 0; off < p.w_bytes; off += 16384)
          bulk_copy_g2s(smW + off, w_src + off, static_cast<uint32_t>(min(16384, p.w_bytes - off)), wbar);
      }
      pdl_wait();   // (weights / bias are constants: loaded above, before the previous kernel has finished)
      int s = 0, bs = 0;
      uint32_t ph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t = tile;
        const int nt = t % p.num_n_tiles; t /= p.num_n_tiles;
        const int ta = t % p.dim_a; t /= p.dim_a;
        const int tb = t % p.dim_b; t /= p.dim_b;
        const int tx = p.ty_first ? tb : ta, ty = p.ty_first ? ta : tb;
        const int tz = t % p.tiles_z;
        const int n = t / p.tiles_z;
        const int cx = tx * p.x_step + p.x_off, cy = ty * p.y_step + p.y_off, cz = tz * p.z_step + p.z_off;
        for (int ks = 0; ks < p.k_stages; ++ks) {
          const int kt = ks / p.cb_n, cb = ks - kt * p.cb_n;   // K stage = (temporal tap, 64-channel block)
          mbar_wait(empty + s, ph ^ 1);
          if (HAS_UP && cb >= p.up_cb_first) {
            mbar_arrive(full + s);   // this slab is written by the up-sampling warps
          } else if (PAIR) {
            // both slabs complete on the LEADER's barrier (its issuer reads both shared memories)
            if (rank == 0) mbar_arrive_expect_tx(full + s, 2u * static_cast<uint32_t>(p.slab_bytes));
            if (p.merged_cw)   // stems (STEM3D_PAIR): (pixel, channel) merged into one contiguous inner dimension
              tma_load_5d_pair(smS + s * p.slab_stride, &p.tmA, mapa_u32(smem_u32(full + s), 0), cx * 8, cy,
                               cz + kt * p.z_kstep, n, 0);
            else
              tma_load_5d_pair(smS + s * p.slab_stride, &p.tmA, mapa_u32(smem_u32(full + s), 0), cb * p.c_step, cx, cy,
                               cz + kt * p.z_kstep, n);
          } else {
            mbar_arrive_expect_tx(full + s, static_cast<uint32_t>(p.slab_bytes));
            if (p.merged_cw)  // stems: (pixel, channel) merged into one contiguous inner dimension of 8-element pixels
              tma_load_5d(smS + s * p.slab_stride, &p.tmA, full + s, cx * 8, cy, cz + kt * p.z_kstep, n, 0);
            else
              tma_load_5d(smS + s * p.slab_stride, &p.tmA, full + s, cb * p.c_step, cx, cy, cz + kt * p.z_kstep, n);
          }
          if (++s == S) { s = 0; ph ^= 1; }
          if (p.b_stream) {
            // one [n_tile x 64] weight block per filter tap, K ordered (kd,kh,kw,cin): tap kt*n_grp+g, block cb
            for (int g = 0; g < p.n_grp; ++g) {
              mbar_wait(bempty + bs, bph ^ 1);
              if (PAIR) {
                // each CTA streams HALF of the block's rows; both halves complete on the leader's barrier
                if (rank == 0) mbar_arrive_expect_tx(bfull + bs, 2u * static_cast<uint32_t>(p.b_stride));
                tma_load_2d_pair(smW + bs * p.b_stride, &p.tmB, mapa_u32(smem_u32(bfull + bs), 0),
                                 (kt * p.n_grp + g) * p.cin + cb * 64, nt * p.n_tile + static_cast<int>(rank) * (p.n_tile >> 1));
              } else {
                mbar_arrive_expect_tx(bfull + bs, static_cast<uint32_t>(p.b_stride));
                tma_load_2d(smW + bs * p.b_stride, &p.tmB, bfull + bs, (kt * p.n_grp + g) * p.cin + cb * 64, nt * p.n_tile);
              }
              if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // -------------------------------------------------------------- MMA issuer
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t wa = smem_u32(smW);
#define TSP_ISSUE(TM_, NK_, ST_) slab_issue<TM_, NK_, ST_, false>(p, smS, wa, tb, full, empty, tfull, tempty, wbar, bfull, bempty)
    if (PAIR) {
      if (rank == 0) {
#define TSP_ISSUE_PAIR(TM_, ST_) slab_issue<TM_, 4, ST_, true>(p, smS, wa, tb, full, empty, tfull, tempty, wbar, bfull, bempty)
        if (p.b_stream) {
          if (p.tm == 2) TSP_ISSUE_PAIR(2, true); else TSP_ISSUE_PAIR(1, true);
        } else if (p.kx) {        // KX kind: one 128-row tile per CTA, three filter-row groups per K stage
          slab_issue<1, 4, false, true>(p, smS, wa, tb, full, empty, tfull, tempty, wbar, bfull, bempty);
        } else if (p.nk == 2) {   // un-swizzled stem (STEM3D_PAIR): two K = 16 steps per filter row
          if (p.tm == 2) slab_issue<2, 2, false, true>(p, smS, wa, tb, full, empty, tfull, tempty, wbar, bfull, bempty);
          else slab_issue<1, 2, false, true>(p, smS, wa, tb, full, empty, tfull, tempty, wbar, bfull, bempty);
        } else {
          TSP_ISSUE_PAIR(2, false);
        }
#undef TSP_ISSUE_PAIR
      } else if (!p.b_stream) {
        // peer: report its half of the weights resident, then leave the issuing to the leader
        mbar_wait(wbar, 0);
        if (lane == 0) mbar_arrive_cluster_release(mapa_u32(smem_u32(bfull), 0));
      }
    } else if (p.b_stream) {
      if (p.tm == 2) TSP_ISSUE(2, 4, true); else TSP_ISSUE(1, 4, true);
    } else if (p.tm == 2) {
      if (p.nk == 4) TSP_ISSUE(2, 4, false); else TSP_ISSUE(2, 2, false);
    } else {
      if (p.nk == 4) TSP_ISSUE(1, 4, false); else TSP_ISSUE(1, 2, false);
    }
#undef TSP_ISSUE
  } else if (warp < NEW) {
    // ---------------------------------------------------------------- epilogue (slab_epilogue above)
    const bool relu = p.act == TEDSPAD_ACT_RELU;
    if (!E16 && !HAS_UP && p.kx) {
      if (relu) slab_epilogue_kx<PAIR, true>(p, warp, lane, tmem_base, sm_bias, tfull, tempty);
      else slab_epilogue_kx<PAIR, false>(p, warp, lane, tmem_base, sm_bias, tfull, tempty);
    } else {
#define TSP_EPI(MODE_) \
    do { \
      if (relu) slab_epilogue<HAS_UP, PAIR, MODE_, true, E16>(p, warp, lane, tmem_base, sm_bias, sm_ocw, tfull, tempty, sm_part, smem + p.tst_off); \
      else slab_epilogue<HAS_UP, PAIR, MODE_, false, E16>(p, warp, lane, tmem_base, sm_bias, sm_ocw, tfull, tempty, sm_part, smem + p.tst_off); \
    } while (0)
    if (p.res != nullptr) TSP_EPI(EPI_RES);
    else if (p.oc_w != nullptr && p.y == nullptr && p.pool == nullptr) TSP_EPI(EPI_OC_ONLY);
    else if (p.oc_w != nullptr) TSP_EPI(EPI_OC);
    else if (p.pool != nullptr) TSP_EPI(EPI_POOL);
    else TSP_EPI(EPI_PLAIN);
#undef TSP_EPI
    }
  }

  if (HAS_UP && warp >= 11) {
    // -------------------------------------------------- up-sampling slab producers (warps 11-14)
    // thread -> one 16-byte channel chunk (8 channels) of 16 slab pixels per pass; the slab row of pixel r is
    // r*128 bytes with the SWIZZLE_128B chunk permutation the TMA would have applied.
    pdl_wait();
    const int it = static_cast<int>(threadIdx.x) - 11 * 32;
    const int c8 = it & 7, p0 = it >> 3;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int t = tile;
      t /= p.num_n_tiles;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int n = t / p.tiles_z;
      const int iy0 = ty * p.y_step - 1, ix0 = tx * p.x_step - 1;   // image coordinates of slab pixel (0,0)
      for (int ks = 0; ks < p.k_stages; ++ks) {
        const int cb = ks % p.cb_n;
        mbar_wait(empty + s, ph ^ 1);
        if (cb >= p.up_cb_first) {
          uint8_t* slab = smS + s * p.slab_stride;
          const int ch = (cb - p.up_cb_first) * 64 + c8 * 8 + p.up_coff;
          // software pipeline: the four taps of pixel r+16 are in flight while pixel r is interpolated
          UpTaps cur = up_fetch(p, p0, iy0, ix0, n, ch);
          for (int r = p0; r < p.slab_px; r += 16) {
            const UpTaps nxt = (r + 16 < p.slab_px) ? up_fetch(p, r + 16, iy0, ix0, n, ch) : cur;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (cur.in) {
              const uint32_t* pa = reinterpret_cast<const uint32_t*>(&cur.qa);
              const uint32_t* pb = reinterpret_cast<const uint32_t*>(&cur.qb);
              const uint32_t* pc = reinterpret_cast<const uint32_t*>(&cur.qc);
              const uint32_t* pd = reinterpret_cast<const uint32_t*>(&cur.qd);
              uint32_t* po = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float a0 = __uint_as_float(pa[i] << 16), a1 = __uint_as_float(pa[i] & 0xffff0000u);
                const float b0 = __uint_as_float(pb[i] << 16), b1 = __uint_as_float(pb[i] & 0xffff0000u);
                const float c0 = __uint_as_float(pc[i] << 16), c1 = __uint_as_float(pc[i] & 0xffff0000u);
                const float d0 = __uint_as_float(pd[i] << 16), d1 = __uint_as_float(pd[i] & 0xffff0000u);
                // separable form, the same operation order as upsample2x_kernel (ops.cu)
                const float o0 = fmaf(cur.ly1, fmaf(cur.lx1, d0, cur.lx0 * c0), cur.ly0 * fmaf(cur.lx1, b0, cur.lx0 * a0));
                const float o1 = fmaf(cur.ly1, fmaf(cur.lx1, d1, cur.lx0 * c1), cur.ly0 * fmaf(cur.lx1, b1, cur.lx0 * a1));
                po[i] = cvt_bf16x2(o0, o1, false);
              }
            }
            *reinterpret_cast<uint4*>(slab + r * 128 + ((c8 ^ (r & 7)) << 4)) = out;
            cur = nxt;
          }
          fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full + s);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the leader's MMAs read the peer's shared memory to the end
  tc_fence_after();
  if (warp == W_ALLOC) {
    if (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ weight image
struct PackP {
  const __nv_bfloat16* w_std;
  __nv_bfloat16* image;
  int kind, n_tile, K_pad, cin_pad, kd, kh, kw, shift;
  int kx_cp, kx_row0;   // KX kind: channels per filter-column block, first GEMM column of this CTA's half
  long long total;  // image elements
};

// One thread per image element: where does it come from in the standard [Cout_pad][K_pad] layout?
__global__ void __launch_bounds__(256) slab_pack_kernel(const PackP p) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < p.total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long src = -1;
    int n = 0;
    if (p.kind == TEDSPAD_SLAB_3X3_KX_PAIR) {
      // image of one CTA: [ky*CB + cb] blocks of n_tile (= 3*CP/2) rows x 128 B, SWIZZLE_128B; GEMM column
      // kx_row0 + row = filter column kx = col / CP, output channel col % CP
      const long long byte = idx * 2;
      const int blk_bytes = p.n_tile * 128;
      const int blk = static_cast<int>(byte / blk_bytes);
      const int o = static_cast<int>(byte - static_cast<long long>(blk) * blk_bytes);
      const int grp = o >> 10, row = (o >> 7) & 7, chunk_sw = (o >> 4) & 7, within = (o & 15) >> 1;
      const int col = p.kx_row0 + grp * 8 + row;
      const int kxi = col / p.kx_cp;
      n = col - kxi * p.kx_cp;
      const int k = ((chunk_sw ^ row) << 3) + within;
      const int cin = p.cin_pad, cb_n = cin / 64;
      const int ky = blk / cb_n, cb = blk - ky * cb_n;
      src = static_cast<long long>(ky * 3 + kxi) * cin + cb * 64 + k;
    } else if (p.kind == TEDSPAD_SLAB_3X3) {
      // image: [tap*CB + cb] blocks of n_tile rows x 128 B, SWIZZLE_128B, 8-row groups 1024 B apart
      const long long byte = idx * 2;
      const int blk_bytes = p.n_tile * 128;
      const int blk = static_cast<int>(byte / blk_bytes);
      const int o = static_cast<int>(byte - static_cast<long long>(blk) * blk_bytes);
      const int grp = o >> 10, row = (o >> 7) & 7, chunk_sw = (o >> 4) & 7, within = (o & 15) >> 1;
      n = grp * 8 + row;
      const int k = ((chunk_sw ^ row) << 3) + within;
      const int cin = p.cin_pad, cb_n = cin / 64;
      const int tap = blk / cb_n, cb = blk - tap * cb_n;
      src = static_cast<long long>(tap) * cin + cb * 64 + k;
    } else {
      // image: [mma][chunk j (2)][n][8 elements], un-swizzled
      long long t = idx;
      const int e = static_cast<int>(t & 7); t >>= 3;
      n = static_cast<int>(t % p.n_tile); t /= p.n_tile;
      const int j = static_cast<int>(t & 1); t >>= 1;
      const int mma = static_cast<int>(t);
      if (p.kind == TEDSPAD_SLAB_STEM2D) {
        // mma = ky*2 + q; chunk = tap kx = 2q + j (kx == 3: zero), 8 channels of one pixel
        const int ky = mma >> 1, q = mma & 1, kx = 2 * q + j;
        if (kx < 3 && e < p.cin_pad) src = static_cast<long long>(ky * 3 + kx) * p.cin_pad + e;
      } else {
        // mma = (kt*kh + ky)*2 + q; chunk = pixel pair pp = 2q + j -> pixels 2pp, 2pp+1 (4 channels each);
        // filter tap kx = pixel - shift (shift = 1 when the front pad is odd)
        const int q = mma & 1;
        const int kyt = mma >> 1;
        const int ky = kyt % p.kh, kt = kyt / p.kh;
        const int px = 2 * (2 * q + j) + (e >> 2), ch = e & 3;
        const int kx = px - p.shift;
        if (kx >= 0 && kx < p.kw && ch < p.cin_pad)
          src = (static_cast<long long>(kt * p.kh + ky) * p.kw + kx) * p.cin_pad + ch;
      }
    }
    p.image[idx] = src >= 0 ? p.w_std[static_cast<long long>(n) * p.K_pad + src] : __float2bfloat16_rn(0.f);
  }
}

static long long slab_image_bytes(int kind, int n_tile, int cin_pad, int kd, int kh, int kw) {
  if (kind == TEDSPAD_SLAB_3X3) return 9LL * (cin_pad / 64) * n_tile * 128;
  if (kind == TEDSPAD_SLAB_STEM2D) return 6LL * 2 * n_tile * 16;
  return static_cast<long long>(kd) * kh * 2 * 2 * n_tile * 16;
}

// ------------------------------------------------------------------------------------------- plan
// The sixteen-warp epilogue (slab_epilogue, E16): resident-weight CTA-pair layers with 64 outputs and 16x16 tiles.
// TEDSPAD_SLAB_E16=0 falls back to the eight-warp epilogue (same arithmetic, for A/B timing).
static bool plan_e16(const tedspad_slab_plan& P, bool fused_oc) {
  static const int mode = [] {   // 0 = never, 1 = layers with a fused OutConv (default), 2 = every eligible layer
    const char* e = getenv("TEDSPAD_SLAB_E16");
    return e == nullptr ? 1 : atoi(e);
  }();
  return (mode >= 2 || (mode == 1 && fused_oc)) && P.pair && !P.b_stream && P.swizzle128 && P.n_tile == 64 && P.tm == 2;
}

static int make_plan(const tedspad_conv_slab& c, tedspad_slab_plan& P) {
  memset(&P, 0, sizeof(P));
  const tedspad_tensor& x = c.x;
  const tedspad_tensor& y = c.y;
  const bool stream = c.kind == TEDSPAD_SLAB_3X3_STREAM || c.kind == TEDSPAD_SLAB_3X3_STREAM_PAIR;
  const bool kx = c.kind == TEDSPAD_SLAB_3X3_KX_PAIR;
  const bool pair = c.kind == TEDSPAD_SLAB_3X3_PAIR || c.kind == TEDSPAD_SLAB_3X3_STREAM_PAIR || c.kind == TEDSPAD_SLAB_STEM3D_PAIR || kx;
  P.pair = pair ? 1 : 0;
  TSP_CHECK(c.Cout_pad % 32 == 0 && c.Cout_pad >= 32 && c.Cout_pad <= (stream ? 2048 : 256) && c.Cout <= c.Cout_pad &&
                c.Cout >= 1 && c.Cout % 8 == 0,
            "slab: Cout=%d (multiple of 8) / Cout_pad=%d (multiple of 32, <= %d) invalid", c.Cout, c.Cout_pad,
            stream ? 2048 : 256);
  TSP_CHECK(x.N == y.N && x.N >= 1, "slab: batch mismatch");
  P.n_tile = c.Cout_pad;
  P.num_n_tiles = 1;
  P.cb_n = 1;
  P.tab_per_stage = 1;
  if (stream) {
    P.n_tile = c.n_tile > 0 ? c.n_tile : c.Cout_pad / ((c.Cout_pad + 255) / 256);
    TSP_CHECK(P.n_tile % 32 == 0 && P.n_tile <= 256 && c.Cout_pad % P.n_tile == 0, "slab stream: n_tile=%d does not tile Cout_pad=%d",
              P.n_tile, c.Cout_pad);
    P.num_n_tiles = c.Cout_pad / P.n_tile;
  }
  const int Wp = x.W + 2 * x.pw, Hp = x.H + 2 * x.ph, Dp = x.D + 2 * x.pd;
  int slab_w = 0, slab_h = 0, pad_bytes = 0;
  // fused up-sampling: the convolution sees [x | upsample2x(up)] along the channels
  const bool has_up = c.up.ptr != nullptr;
  if (has_up) {
    TSP_CHECK(c.kind == TEDSPAD_SLAB_3X3 || c.kind == TEDSPAD_SLAB_3X3_STREAM, "slab: fused up-sampling needs a single-CTA 3x3 kind");
    TSP_CHECK(c.up.C % 64 == 0 && c.up.C >= 64 && c.up.N == x.N && c.up.D == 1 && x.D == 1 && c.kd == 1 &&
                  2 * c.up.H <= x.H && 2 * c.up.W <= x.W,
              "slab: up-sampling source [%d,%d,%d,%d] does not fit the [%d,%d,%d] input", c.up.N, c.up.H, c.up.W, c.up.C,
              x.N, x.H, x.W);
  }
  const int cin_total = x.C + (has_up ? c.up.C : 0);
  P.up_cb_first = has_up ? x.C / 64 : 1 << 20;
  if (kx) {
    // three filter columns stacked along N (slab_epilogue_kx): tile = 8 rows x 16 slab columns -> 14 output columns
    TSP_CHECK(c.kd == 1 && c.kh == 3 && c.kw == 3 && c.sd == 1 && c.sh == 1 && c.sw == 1 && c.pd == 0 && c.ph == 1 && c.pw == 1,
              "slab kx: needs a (1,3,3) stride-1 pad-1 convolution");
    TSP_CHECK(x.D == 1 && y.D == 1 && x.H == y.H && x.W == y.W, "slab kx: 2-D layers with equal input / output extents only");
    TSP_CHECK(x.C % 64 == 0 && x.C >= 64 && !has_up, "slab kx: x.C=%d must be a multiple of 64 (no fused up-sampling)", x.C);
    TSP_CHECK(c.Cout_pad == 32 || c.Cout_pad == 64, "slab kx: Cout_pad=%d must be 32 or 64", c.Cout_pad);
    TSP_CHECK(c.pool.ptr == nullptr && c.oc_w == nullptr && c.res == nullptr, "slab kx: no fused pool / OutConv / residual");
    TSP_CHECK(c.oc_clip.ptr == nullptr || (c.Cout_pad == 32 && c.Cout >= 12),
              "slab kx: the space-to-depth clip glue needs the 12-channel head (Cout_pad 32), got Cout=%d", c.Cout);
    const int cp = c.Cout_pad, cb_n = x.C / 64;
    P.n_tile = 3 * cp;
    const int b_rows = P.n_tile / 2;                       // weight rows held by one CTA of the pair
    P.w_bytes = 3 * cb_n * b_rows * 128;
    P.swizzle128 = 1;
    P.k_stages = cb_n; P.cb_n = cb_n;
    P.n_grp = 3; P.nk = 4; P.n_mma = 12;
    P.tm = 1;
    slab_w = 16; slab_h = 10;
    P.box[0] = 64; P.box[1] = slab_w; P.box[2] = slab_h; P.box[3] = 1; P.box[4] = 1;
    P.tdim[0] = x.C; P.tdim[1] = Wp; P.tdim[2] = Hp; P.tdim[3] = Dp; P.tdim[4] = x.N;
    P.tstride[0] = static_cast<int64_t>(x.ld) * 2;
    P.tstride[1] = P.tstride[0] * Wp;
    P.tstride[2] = P.tstride[1] * Hp;
    P.tstride[3] = P.tstride[2] * Dp;
    P.tbase_off = static_cast<int64_t>(x.coff) * 2;
    P.c_step = 64;
    P.x_step = 14; P.x_off = x.pw - 1;
    P.y_step = 8; P.y_off = x.ph - 1;
    P.z_step = 1; P.z_off = x.pd; P.z_kstep = 0;
    P.tiles_x = (x.W + 13) / 14;
    P.tiles_y = (x.H + 7) / 8;
    P.tiles_z = 1;
    P.half_a_off = 0;
    P.a_layout = 2; P.a_lbo = 16; P.a_sbo = 1024;   // the slab is 16 pixels wide: consecutive 8-pixel groups are 1024 B apart
    P.b_layout = 2; P.b_lbo = 16; P.b_sbo = 1024;
    P.a_kstep = 32; P.b_kstep = 32;
    for (int cb = 0; cb < cb_n; ++cb)
      for (int ky = 0; ky < 3; ++ky) {
        const int i = cb * 3 + ky;
        TSP_CHECK(i < TEDSPAD_SLAB_MAX_MMA, "slab kx: table overflow");
        P.tab[2 * i] = static_cast<uint32_t>(ky * slab_w * 128);
        P.tab[2 * i + 1] = static_cast<uint32_t>((ky * cb_n + cb) * b_rows * 128);
      }
  } else if (c.kind == TEDSPAD_SLAB_3X3 || c.kind == TEDSPAD_SLAB_STEM2D || c.kind == TEDSPAD_SLAB_3X3_PAIR) {
    TSP_CHECK(c.kd == 1 && c.kh == 3 && c.kw == 3 && c.sd == 1 && c.sh == 1 && c.sw == 1 && c.pd == 0 && c.ph == 1 &&
                  c.pw == 1,
              "slab: kind %d needs a (1,3,3) stride-1 pad-1 convolution", c.kind);
    TSP_CHECK(x.D == y.D && x.H == y.H && x.W == y.W, "slab: output extents must equal input extents");
    const bool sw = c.kind == TEDSPAD_SLAB_3X3 || pair;
    const int b_rows = pair ? P.n_tile / 2 : P.n_tile;   // weight rows (output channels) held by one CTA
    if (pair) TSP_CHECK(P.n_tile % 32 == 0 && x.W > 8, "slab pair: needs Cout_pad %% 32 == 0 and W > 8 (got %d, %d)", P.n_tile, x.W);
    if (sw) {
      TSP_CHECK(x.C % 64 == 0 && x.C >= 64, "slab 3x3: x.C=%d must be a multiple of 64", x.C);
      // no halo needed: taps outside the tensor are zero-filled by TMA (a zero halo works just as well)
    } else {
      TSP_CHECK(x.C == 8, "slab stem2d: x.C=%d must be 8 (channels padded to one 16-byte pixel)", x.C);
    }
    P.w_bytes = static_cast<int>(slab_image_bytes(sw ? TEDSPAD_SLAB_3X3 : c.kind, b_rows, sw ? cin_total : x.C, 1, 3, 3));
    P.swizzle128 = sw ? 1 : 0;
    P.k_stages = sw ? cin_total / 64 : 1;
    P.cb_n = P.k_stages;
    P.n_grp = sw ? 9 : 3;     // filter taps (3x3) / filter rows (stem)
    P.nk = sw ? 4 : 2;        // K=16 steps per group
    P.n_mma = P.n_grp * P.nk;
    const int px_bytes = sw ? 128 : 16;
    // tile width: two 8-column halves when the image is wide enough and three slab stages still fit
    int tm = c.tm;
    if (tm == 0) {
      tm = x.W > 8 ? 2 : 1;
      if (tm == 2) {
        const int stride2 = static_cast<int>(round_up(px_bytes * 18 * 18 + (sw ? 0 : 64), 1024));
        if (P.w_bytes + 3 * stride2 + SLAB_TAIL_BYTES + 1024 > SLAB_SMEM_BUDGET) tm = 1;
      }
    }
    TSP_CHECK(tm == 1 || tm == 2, "slab: tm=%d", tm);
    TSP_CHECK(!pair || tm == 2, "slab pair: the tile must be two 8-column halves (weights too large for three slab stages?)");
    P.tm = tm;
    slab_w = 8 * tm + 2;
    slab_h = 18;
    pad_bytes = sw ? 0 : 64;  // stem: the zero-weight tap kx=3 of the last row reads past the box
    if (sw) {
      P.box[0] = 64; P.box[1] = slab_w; P.box[2] = slab_h; P.box[3] = 1; P.box[4] = 1;
      P.tdim[0] = x.C; P.tdim[1] = Wp; P.tdim[2] = Hp; P.tdim[3] = Dp; P.tdim[4] = x.N;
      P.tstride[0] = static_cast<int64_t>(x.ld) * 2;
      P.tstride[1] = P.tstride[0] * Wp;
      P.tstride[2] = P.tstride[1] * Hp;
      P.tstride[3] = P.tstride[2] * Dp;
    } else {
      // pixels are 16 contiguous bytes: merge (pixel, channel) into one inner dimension so that a slab row
      // is ONE contiguous TMA row of slab_w*16 bytes instead of slab_w 16-byte rows
      TSP_CHECK(x.ld == 8 && x.coff == 0, "slab stem2d: input must be a dense 8-channel image (ld=%d coff=%d)", x.ld, x.coff);
      P.merged_cw = 1;
      P.box[0] = slab_w * 8; P.box[1] = slab_h; P.box[2] = 1; P.box[3] = 1; P.box[4] = 1;
      P.tdim[0] = Wp * 8; P.tdim[1] = Hp; P.tdim[2] = Dp; P.tdim[3] = x.N; P.tdim[4] = 1;
      P.tstride[0] = static_cast<int64_t>(Wp) * 16;
      P.tstride[1] = P.tstride[0] * Hp;
      P.tstride[2] = P.tstride[1] * Dp;
      P.tstride[3] = P.tstride[2] * x.N;
    }
    P.tbase_off = static_cast<int64_t>(x.coff) * 2;
    P.c_step = sw ? 64 : 0;
    P.x_step = 8 * tm; P.x_off = x.pw - 1;
    P.y_step = 16; P.y_off = x.ph - 1;
    P.z_step = 1; P.z_off = x.pd; P.z_kstep = 0;
    P.tiles_x = (x.W + 8 * tm - 1) / (8 * tm);
    P.tiles_y = (x.H + 15) / 16;
    P.tiles_z = x.D;
    P.half_a_off = 8 * px_bytes;
    if (sw) {
      P.a_layout = 2; P.a_lbo = 16; P.a_sbo = slab_w * 128;
      P.b_layout = 2; P.b_lbo = 16; P.b_sbo = 1024;
      P.a_kstep = 32; P.b_kstep = 32;
      const int cb_n = cin_total / 64;
      for (int cb = 0; cb < cb_n; ++cb)
        for (int tap = 0; tap < 9; ++tap) {
          const int i = cb * 9 + tap;
          TSP_CHECK(i < TEDSPAD_SLAB_MAX_MMA, "slab 3x3: %d table groups exceed the table (Cin too large)", i + 1);
          P.tab[2 * i] = static_cast<uint32_t>(((tap / 3) * slab_w + (tap % 3)) * 128);
          P.tab[2 * i + 1] = static_cast<uint32_t>((tap * cb_n + cb) * b_rows * 128);
        }
    } else {
      P.a_layout = 0; P.a_lbo = 16; P.a_sbo = slab_w * 16;
      P.b_layout = 0; P.b_lbo = P.n_tile * 16; P.b_sbo = 128;
      P.a_kstep = 32; P.b_kstep = 2 * P.n_tile * 16;   // next K=16 step: two pixels further, next weight block
      for (int ky = 0; ky < 3; ++ky) {
        P.tab[2 * ky] = static_cast<uint32_t>(ky * slab_w * 16);
        P.tab[2 * ky + 1] = static_cast<uint32_t>(ky * 2 * P.b_kstep);
      }
    }
  } else if (stream) {
    const bool k33 = c.kh == 3 && c.kw == 3 && c.ph == 1 && c.pw == 1;
    const bool k11 = c.kh == 1 && c.kw == 1 && c.ph == 0 && c.pw == 0;   // 1x1x1 and (3,1,1): one tap per K stage
    // strided 1x1x1 (the down-sample projections of the ResNet encoders: torchvision resnet.py:96-99, large_i3d.py:
    // 150-156, video/resnet.py): the slab of a tile IS the tile, so a stride is nothing but the traversal stride of the
    // TMA box (cuTensorMapEncodeTiled elementStrides) - the slab holds every sh-th / sw-th pixel of a 2x larger window
    const bool strided = k11 && c.kd == 1 && (c.sd > 1 || c.sh > 1 || c.sw > 1);
    TSP_CHECK((c.kd == 1 || c.kd == 3) && (k33 || k11) && c.pd == c.kd / 2 &&
                  ((c.sd == 1 && c.sh == 1 && c.sw == 1) || (strided && c.sd <= 2 && c.sh <= 2 && c.sw <= 2 && !has_up && !pair)),
              "slab stream: needs a (1|3) x (3x3 | 1x1) stride-1 same-padded convolution, or a 1x1x1 one with strides <= 2");
    const int ntap = c.kh * c.kw, hw = c.kw / 2;   // spatial taps per K stage, halo of the slab
    if (strided)
      TSP_CHECK(y.D == (x.D - 1) / c.sd + 1 && y.H == (x.H - 1) / c.sh + 1 && y.W == (x.W - 1) / c.sw + 1 && c.pool.ptr == nullptr,
                "slab stream: strided 1x1x1 output [%d,%d,%d] does not match input [%d,%d,%d]", y.D, y.H, y.W, x.D, x.H, x.W);
    else
      TSP_CHECK(x.D == y.D && x.H == y.H && x.W == y.W, "slab: output extents must equal input extents");
    TSP_CHECK(x.C % 64 == 0 && x.C >= 64, "slab stream: x.C=%d must be a multiple of 64", x.C);
    TSP_CHECK(c.K_pad == c.kd * ntap * cin_total, "slab stream: K_pad=%d != %d taps x %d channels", c.K_pad, c.kd * ntap, cin_total);
    TSP_CHECK(c.oc_w == nullptr, "slab stream: fused OutConv needs the resident-weight kind");
    P.swizzle128 = 1;
    P.cb_n = cin_total / 64;
    P.cin = cin_total;
    P.k_stages = c.kd * P.cb_n;
    P.n_grp = ntap; P.nk = 4; P.n_mma = 4 * ntap;
    P.tab_per_stage = 0;      // the tap offsets are the same for every K stage
    P.b_stream = 1;
    // CTA pairs: every CTA streams half of each weight block's rows (the MMA reads N/2 rows from either SM), which
    // takes the N = 128 layers off the shared-memory read limit ((128 + 128) rows per K step = exactly 128 B/clk,
    // 70-80 % tensor-active once the TMA fills compete; (128 + 64) rows leave room) and halves the weight traffic
    TSP_CHECK(!pair || (P.num_n_tiles == 1 && P.n_tile % 32 == 0),
              "slab stream pair: one N tile (Cout_pad <= 256, multiple of 32) needed, got %d x %d", P.n_tile, P.num_n_tiles);
    P.b_stride = (pair ? P.n_tile / 2 : P.n_tile) * 128;
    int tm = c.tm;
    if (tm == 0) tm = (x.W > 8 && 4 * P.n_tile <= 512) ? 2 : 1;
    TSP_CHECK((tm == 1 || tm == 2) && 2 * tm * P.n_tile <= 512, "slab stream: tm=%d with n_tile=%d exceeds TMEM", tm, P.n_tile);
    P.tm = tm;
    slab_w = 8 * tm + 2 * hw;
    slab_h = 16 + 2 * hw;
    P.box[0] = 64; P.box[1] = slab_w; P.box[2] = slab_h; P.box[3] = 1; P.box[4] = 1;
    P.tdim[0] = x.C; P.tdim[1] = Wp; P.tdim[2] = Hp; P.tdim[3] = Dp; P.tdim[4] = x.N;
    P.tstride[0] = static_cast<int64_t>(x.ld) * 2;
    P.tstride[1] = P.tstride[0] * Wp;
    P.tstride[2] = P.tstride[1] * Hp;
    P.tstride[3] = P.tstride[2] * Dp;
    P.tbase_off = static_cast<int64_t>(x.coff) * 2;
    // no halo needed: taps outside the tensor are zero-filled by TMA (a zero halo works just as well)
    P.c_step = 64;
    // (x_step / y_step / z_step are in INPUT pixels: a strided tile starts sw * 8 * tm columns after its neighbour)
    P.x_step = 8 * tm * c.sw; P.x_off = x.pw - hw;
    P.y_step = 16 * c.sh; P.y_off = x.ph - hw;
    P.z_step = c.sd; P.z_off = x.pd - c.kd / 2; P.z_kstep = 1;
    P.tiles_x = (y.W + 8 * tm - 1) / (8 * tm);
    P.tiles_y = (y.H + 15) / 16;
    P.tiles_z = y.D;
    P.half_a_off = 8 * 128;
    P.a_layout = 2; P.a_lbo = 16; P.a_sbo = slab_w * 128;
    P.b_layout = 2; P.b_lbo = 16; P.b_sbo = 1024;
    P.a_kstep = 32; P.b_kstep = 32;
    for (int tap = 0; tap < ntap; ++tap) {
      P.tab[2 * tap] = static_cast<uint32_t>(((tap / c.kw) * slab_w + (tap % c.kw)) * 128);
      P.tab[2 * tap + 1] = 0;
    }
  } else if (c.kind == TEDSPAD_SLAB_STEM3D || c.kind == TEDSPAD_SLAB_STEM3D_PAIR) {
    TSP_CHECK(c.kh == 7 && c.kw == 7 && c.sh == 2 && c.sw == 2 && c.kd >= 1 && c.kd <= 7 && c.sd >= 1,
              "slab stem3d: needs a (kd,7,7) stride (sd,2,2) convolution");
    TSP_CHECK(x.C == 4 && x.ld == 4 && x.coff == 0 && x.pd == 0 && x.ph == 0 && x.pw == 0 && x.W % 2 == 0,
              "slab stem3d: input must be an un-haloed [N][D][H][W][4] clip with even W");
    TSP_CHECK(c.pw >= 0 && c.pw <= 6 && c.ph >= 0 && c.ph <= 6 && c.pd >= 0 && c.pd < c.kd, "slab stem3d: bad front pads");
    const int shift = c.pw & 1;          // odd front pad: one extra (zero-weight) leading tap
    const int pwe = c.pw + shift;        // even front pad in pixels
    TSP_CHECK(7 + shift <= 8, "slab stem3d: window exceeds 8 pixels");
    const int b_rows = pair ? P.n_tile / 2 : P.n_tile;   // weight rows (output channels) held by one CTA
    TSP_CHECK(!pair || P.n_tile % 32 == 0, "slab stem3d pair: Cout_pad=%d must be a multiple of 32", P.n_tile);
    P.w_bytes = static_cast<int>(slab_image_bytes(TEDSPAD_SLAB_STEM3D, b_rows, 4, c.kd, 7, 7));
    int tm = c.tm;
    if (tm == 0) tm = 1;
    TSP_CHECK(tm == 1 || tm == 2, "slab: tm=%d", tm);
    P.tm = tm;
    const int pairs = 8 * tm + 3;        // 16*tm + 6 pixels
    slab_w = pairs;
    slab_h = 37;
    pad_bytes = 64;
    P.k_stages = c.kd;
    P.n_grp = 7;   // filter rows
    P.nk = 2;      // two K=16 steps = 4 pixel pairs = the 8-pixel window of one filter row
    P.n_mma = 14;
    // (pixel pair, channel) merged into one contiguous inner dimension: a slab row is one TMA row
    P.merged_cw = 1;
    P.box[0] = pairs * 8; P.box[1] = slab_h; P.box[2] = 1; P.box[3] = 1; P.box[4] = 1;
    P.tdim[0] = x.W * 4; P.tdim[1] = x.H; P.tdim[2] = x.D; P.tdim[3] = x.N; P.tdim[4] = 1;
    P.tstride[0] = static_cast<int64_t>(x.W) * 8;
    P.tstride[1] = P.tstride[0] * x.H;
    P.tstride[2] = P.tstride[1] * x.D;
    P.tstride[3] = P.tstride[2] * x.N;
    P.tbase_off = 0;
    P.c_step = 0;
    P.x_step = 8 * tm; P.x_off = -(pwe / 2);
    P.y_step = 32; P.y_off = -c.ph;
    P.z_step = c.sd; P.z_off = -c.pd; P.z_kstep = 1;
    P.tiles_x = (y.W + 8 * tm - 1) / (8 * tm);
    P.tiles_y = (y.H + 15) / 16;
    P.tiles_z = y.D;
    P.half_a_off = 8 * 16;
    P.a_layout = 0; P.a_lbo = 16; P.a_sbo = 2 * pairs * 16;
    P.b_layout = 0; P.b_lbo = b_rows * 16; P.b_sbo = 128;
    P.a_kstep = 32; P.b_kstep = 2 * b_rows * 16;
    for (int kt = 0; kt < c.kd; ++kt)
      for (int ky = 0; ky < 7; ++ky) {
        const int e = kt * 7 + ky;
        TSP_CHECK(e < TEDSPAD_SLAB_MAX_MMA, "slab stem3d: table overflow");
        P.tab[2 * e] = static_cast<uint32_t>(ky * pairs * 16);
        P.tab[2 * e + 1] = static_cast<uint32_t>(e * 2 * P.b_kstep);
      }
    // the last output row/column must read inside the zero-filled box: guaranteed by TMA OOB fill
  } else {
    TSP_CHECK(false, "slab: unknown kind %d", c.kind);
  }
  // Stacked rows.  A 16-row tile grid wastes ceil(H/16)*16/H of the tensor-core time per image (56, 28: +14 %).
  // When the input buffer carries zero halo ROWS (x.ph >= 1, the contract of the anonymizer's buffers), the rows of
  // all images are tiled as ONE column of N*(H+2ph) rows instead: a tile's slab then simply runs across the halo
  // rows between two images (which are the zero padding both need), the per-image remainder disappears and only
  // the halo rows themselves are computed in vain (58/56, 30/28).  Tile row 0 = stacked row ph, so that with an
  // even padded height the 2x2 pooling pairs of the fused MaxPool2d never straddle tiles or images.
  int64_t batch = x.N;
  const bool kind3x3 = c.kind == TEDSPAD_SLAB_3X3 || stream || c.kind == TEDSPAD_SLAB_3X3_PAIR || kx;
  if (kind3x3 && c.stack_rows >= 0 && x.D == 1 && x.pd == 0 && c.kd == 1 && x.ph >= 1 && !has_up && x.N > 1 && c.sh == 1 && c.sw == 1 &&
      (c.stack_rows > 0 || Hp < round_up(x.H, P.y_step)) && (c.pool.ptr == nullptr || Hp % 2 == 0)) {
    P.stack_hp = Hp; P.stack_ph = x.ph; P.stack_n = x.N;
    P.tdim[2] = Hp * x.N; P.tdim[3] = 1; P.tdim[4] = 1;
    P.tstride[2] = P.tstride[1] * Hp * x.N;
    P.tstride[3] = P.tstride[2];
    P.tiles_y = (Hp * x.N - 2 * x.ph + P.y_step - 1) / P.y_step;
    batch = 1;
  }
  // the bias of every N tile lives in shared memory: 512 floats by default, more for the 1024 / 2048-output convolutions
  const int bias_floats = std::max(512, c.Cout_pad);
  const int tail_bytes = SLAB_TAIL_BYTES + (bias_floats - 512) * 4 + (plan_e16(P, c.oc_w != nullptr) ? SLAB_E16_SCRATCH : 0);
  P.slab_bytes = 2 * P.box[0] * P.box[1] * P.box[2] * P.box[3] * P.box[4];
  P.slab_stride = static_cast<int>(round_up(P.slab_bytes + pad_bytes, 1024));
  int w_stride = static_cast<int>(round_up(P.w_bytes, 1024));
  if (P.b_stream) {
    // two (tm=2) or three slab stages; the rest of shared memory is the weight-block ring
    P.stages = (P.tm == 2 && !pair) ? 2 : 3;   // (a pair's half-size weight blocks leave room for a third 16x16 slab)
    const int avail_b = SLAB_SMEM_BUDGET - 1024 - tail_bytes - P.stages * P.slab_stride;
    P.b_stages = std::min(SLAB_MAX_BSTAGES, avail_b / P.b_stride);
    TSP_CHECK(P.b_stages >= 3, "slab stream: only %d weight-block stages fit", P.b_stages);
    w_stride = P.b_stages * P.b_stride;
  } else {
    const int avail = SLAB_SMEM_BUDGET - 1024 - tail_bytes - w_stride;
    P.stages = std::min(SLAB_MAX_STAGES, avail / P.slab_stride);
    TSP_CHECK(P.stages >= 2, "slab: weights (%d B) + two slab stages (%d B each) do not fit in shared memory", P.w_bytes,
              P.slab_stride);
  }
  P.smem_bytes = 1024 + w_stride + P.stages * P.slab_stride + tail_bytes;
  int tc = 32;
  // accumulator ring: the epilogue of tile i overlaps the MMAs of tiles i+1 .. i+acc-1.  Two stages are enough when
  // a tile holds several K stages; with ONE K stage per tile (64 -> 64) the hand-back round trip (commit -> epilogue
  // -> arrive, across the CTA pair) is longer than a tile's MMAs and the issuer stalled on `tempty` (28 polls per
  // tile in profiles/r1d): use every TMEM column that is there.
  P.acc_stages = std::max(2, std::min(SLAB_MAX_ACC, 512 / (P.tm * P.n_tile)));
  while (tc < P.acc_stages * P.tm * P.n_tile) tc <<= 1;
  TSP_CHECK(tc <= 512, "slab: %d TMEM columns needed", tc);
  P.tmem_cols = tc;
  const int64_t total = batch * P.tiles_z * P.tiles_y * P.tiles_x * P.num_n_tiles;
  TSP_CHECK(total > 0 && total < (int64_t(1) << 31), "slab: tile count out of range");
  P.total_tiles = static_cast<int>(total);
  TSP_CHECK(!pair || total % 2 == 0, "slab pair: %lld tiles cannot be split over CTA pairs", (long long)total);
  return 0;
}

}  // namespace tsp

using namespace tsp;

extern "C" int tedspad_conv_slab_plan(const tedspad_conv_slab* c, tedspad_slab_plan* out) {
  TSP_CHECK(c != nullptr && out != nullptr, "slab plan: null argument");
  return make_plan(*c, *out);
}

extern "C" int tedspad_conv_slab_pack(int32_t kind, const void* w_std, int32_t Cout_pad, int32_t K_pad, int32_t cin_pad,
                                      int32_t kd, int32_t kh, int32_t kw, int32_t pw_front, void* image,
                                      int64_t* image_bytes, void* stream) {
  if (kind == TEDSPAD_SLAB_3X3_KX_PAIR) {
    TSP_CHECK((Cout_pad == 32 || Cout_pad == 64) && cin_pad % 64 == 0 && kd == 1 && kh == 3 && kw == 3 && K_pad >= 9 * cin_pad,
              "slab pack kx: bad geometry (Cout_pad=%d cin_pad=%d)", Cout_pad, cin_pad);
    const int rows = 3 * Cout_pad / 2;   // GEMM columns (weight rows) per CTA
    const long long half_bytes = 3LL * (cin_pad / 64) * rows * 128;
    if (image_bytes) *image_bytes = 2 * half_bytes;
    if (image == nullptr) return 0;
    TSP_CHECK(w_std != nullptr, "slab pack: null weights");
    for (int h = 0; h < 2; ++h) {
      PackP p;
      memset(&p, 0, sizeof(p));
      p.w_std = reinterpret_cast<const __nv_bfloat16*>(w_std);
      p.image = reinterpret_cast<__nv_bfloat16*>(image) + h * (half_bytes / 2);
      p.kind = kind; p.n_tile = rows; p.K_pad = K_pad; p.cin_pad = cin_pad; p.kd = 1; p.kh = 3; p.kw = 3;
      p.kx_cp = Cout_pad; p.kx_row0 = h * rows;
      p.total = half_bytes / 2;
      const int blocks = static_cast<int>(std::min<long long>((p.total + 255) / 256, 4096));
      slab_pack_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
      TSP_CUDA(cudaGetLastError());
    }
    return 0;
  }
  TSP_CHECK((kind >= TEDSPAD_SLAB_3X3 && kind <= TEDSPAD_SLAB_STEM3D) || kind == TEDSPAD_SLAB_3X3_PAIR || kind == TEDSPAD_SLAB_STEM3D_PAIR,
            "slab pack: unknown kind %d", kind);
  const bool pair = kind == TEDSPAD_SLAB_3X3_PAIR || kind == TEDSPAD_SLAB_STEM3D_PAIR;
  if (pair) {
    // two single-CTA images of Cout_pad / 2 rows each, the leader's first
    kind = kind == TEDSPAD_SLAB_3X3_PAIR ? TEDSPAD_SLAB_3X3 : TEDSPAD_SLAB_STEM3D;
    TSP_CHECK(Cout_pad % 32 == 0, "slab pack pair: Cout_pad=%d must be a multiple of 32", Cout_pad);
  }
  const int halves = pair ? 2 : 1, rows = Cout_pad / halves;
  TSP_CHECK(Cout_pad % 16 == 0 && Cout_pad >= 16 && Cout_pad <= 256, "slab pack: Cout_pad=%d", Cout_pad);
  if (kind == TEDSPAD_SLAB_3X3) TSP_CHECK(cin_pad % 64 == 0 && kd == 1 && kh == 3 && kw == 3, "slab pack 3x3: bad geometry");
  if (kind == TEDSPAD_SLAB_STEM2D) TSP_CHECK(cin_pad <= 8 && kd == 1 && kh == 3 && kw == 3, "slab pack stem2d: bad geometry");
  if (kind == TEDSPAD_SLAB_STEM3D) TSP_CHECK(cin_pad >= 1 && kh == 7 && kw == 7 && kd >= 1, "slab pack stem3d: bad geometry");
  TSP_CHECK(K_pad >= kd * kh * kw * cin_pad, "slab pack: K_pad=%d too small", K_pad);
  const long long bytes = slab_image_bytes(kind, rows, kind == TEDSPAD_SLAB_STEM3D ? 4 : cin_pad, kd, kh, kw);
  if (image_bytes) *image_bytes = bytes * halves;
  if (image == nullptr) return 0;
  TSP_CHECK(w_std != nullptr, "slab pack: null weights");
  for (int h = 0; h < halves; ++h) {
    PackP p;
    p.w_std = reinterpret_cast<const __nv_bfloat16*>(w_std) + static_cast<long long>(h) * rows * K_pad;
    p.image = reinterpret_cast<__nv_bfloat16*>(image) + h * (bytes / 2);
    p.kind = kind; p.n_tile = rows; p.K_pad = K_pad; p.cin_pad = cin_pad; p.kd = kd; p.kh = kh; p.kw = kw;
    p.shift = pw_front & 1;
    p.total = bytes / 2;
    const int blocks = static_cast<int>(std::min<long long>((p.total + 255) / 256, 4096));
    slab_pack_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    TSP_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int tedspad_conv_slab_forward(const tedspad_conv_slab* c, void* stream_v) {
  TSP_CHECK(c != nullptr, "slab: null descriptor");
  const tedspad_tensor& x = c->x;
  const tedspad_tensor& y = c->y;
  if (check_tensor(x, "slab.x", (c->kind == TEDSPAD_SLAB_STEM3D || c->kind == TEDSPAD_SLAB_STEM3D_PAIR) ? 4 : 8)) return 1;
  const bool fused_oc = c->oc_w != nullptr;
  const bool s2d_clip = c->kind == TEDSPAD_SLAB_3X3_KX_PAIR && !fused_oc && c->oc_clip.ptr != nullptr;
  tedspad_tensor ychk = y;
  if (y.ptr == nullptr) {
    TSP_CHECK(fused_oc || s2d_clip, "slab: y.ptr is NULL without a fused OutConv / clip glue");
    ychk.ptr = const_cast<void*>(c->w_image);  // extents are still validated
  }
  if (check_tensor(ychk, "slab.y", 8)) return 1;
  TSP_CHECK(c->w_image && c->bias, "slab: null weights/bias");
  TSP_CHECK(y.C == c->Cout, "slab: y.C %d != Cout %d", y.C, c->Cout);
  TSP_CHECK(c->act == TEDSPAD_ACT_RELU || c->act == TEDSPAD_ACT_NONE, "slab: activation %d not supported", c->act);
  tedspad_slab_plan P;
  if (int rc = make_plan(*c, P)) return rc;

  SlabKParams p;
  memset(&p, 0, sizeof(p));
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  for (int i = 0; i < 5; ++i) { dims[i] = (uint64_t)P.tdim[i]; box[i] = (uint32_t)P.box[i]; }
  for (int i = 0; i < 4; ++i) strides[i] = (uint64_t)P.tstride[i];
  const uint8_t* base = reinterpret_cast<const uint8_t*>(x.ptr) + P.tbase_off;
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (P.b_stream && (c->sd > 1 || c->sh > 1 || c->sw > 1)) {
    // strided 1x1x1: the box spans s x as many input pixels, the TMA keeps every s-th one (box[] counts what it keeps);
    // one tile covers ONE output depth, so the depth stride is in the tile origin (z_step), not in the box
    estr[1] = (uint32_t)c->sw; estr[2] = (uint32_t)c->sh;
    box[1] = box[1] * estr[1]; box[2] = box[2] * estr[2];
  }
  if (encode_tmap_5d_bf16(&p.tmA, base, dims, strides, box, P.swizzle128 != 0, estr)) return 3;
  p.w_image = reinterpret_cast<const uint8_t*>(c->w_image);
  p.bias = c->bias;
  p.b_stream = P.b_stream; p.b_stages = P.b_stages; p.b_stride = P.b_stride; p.cb_n = P.cb_n; p.cin = P.cin;
  p.num_n_tiles = P.num_n_tiles; p.tab_per_stage = P.tab_per_stage;
  if (P.b_stream) {
    // w_image holds the STANDARD packed weights [Cout_pad][K_pad] for the streaming kind
    if (encode_tmap_2d_bf16(&p.tmB, c->w_image, (uint64_t)c->K_pad, (uint64_t)c->Cout_pad, (uint64_t)c->K_pad * 2, 64,
                            (uint32_t)(P.pair ? P.n_tile / 2 : P.n_tile)))
      return 3;
  }
  p.tm = P.tm; p.n_tile = P.n_tile; p.k_stages = P.k_stages; p.n_grp = P.n_grp; p.nk = P.nk; p.stages = P.stages;
  p.a_kstep = P.a_kstep; p.b_kstep = P.b_kstep;
  p.tmem_cols = P.tmem_cols; p.acc_stages = P.acc_stages;
  p.bias_floats = std::max(512, (int)c->Cout_pad);
  p.slab_bytes = P.slab_bytes; p.slab_stride = P.slab_stride; p.w_bytes = P.w_bytes;
  p.w_stride = P.b_stream ? P.b_stages * P.b_stride : (int)round_up(P.w_bytes, 1024);
  p.zero_slabs = P.swizzle128 ? 0 : 1;
  p.half_a_off = P.half_a_off;
  p.c_step = P.c_step; p.x_step = P.x_step; p.x_off = P.x_off; p.y_step = P.y_step; p.y_off = P.y_off;
  p.z_step = P.z_step; p.z_off = P.z_off; p.z_kstep = P.z_kstep; p.merged_cw = P.merged_cw;
  p.tiles_x = P.tiles_x; p.tiles_y = P.tiles_y; p.tiles_z = P.tiles_z; p.total_tiles = P.total_tiles;
  p.stack_hp = P.stack_hp; p.stack_ph = P.stack_ph; p.stack_n = P.stack_n;
  // CTA pairs with a last column block whose second half is empty (OW % 16 in 1..8): column-major tile order
  p.ty_first = (P.pair && P.tm == 2 && ((y.W - 1) % 16) < 8) ? 1 : 0;
  p.dim_a = p.ty_first ? P.tiles_y : P.tiles_x;
  p.dim_b = p.ty_first ? P.tiles_x : P.tiles_y;
  p.dv_nt = make_div(P.num_n_tiles); p.dv_tx = make_div(p.dim_a); p.dv_ty = make_div(p.dim_b);
  p.kx = c->kind == TEDSPAD_SLAB_3X3_KX_PAIR ? 1 : 0;
  p.kx_cp = c->Cout_pad;
  p.bias_n = p.kx ? c->Cout_pad : P.n_tile * P.num_n_tiles;
  p.dv_tz = make_div(P.tiles_z); p.dv_hp = make_div(P.stack_hp > 0 ? P.stack_hp : 1);
  p.a_desc = umma_desc_template(P.a_layout, P.a_lbo, P.a_sbo);
  p.b_desc = umma_desc_template(P.b_layout, P.b_lbo, P.b_sbo);
  const int n_tab = P.tab_per_stage ? P.k_stages * P.n_grp : P.n_grp;
  for (int i = 0; i < n_tab; ++i) p.tab[i] = make_uint2(P.tab[2 * i] >> 4, P.tab[2 * i + 1] >> 4);  // 16-byte units

  p.y = reinterpret_cast<__nv_bfloat16*>(y.ptr);
  p.OH = y.H; p.OW = y.W;
  p.yDp = y.D + 2 * y.pd; p.yHp = y.H + 2 * y.ph; p.yWp = y.W + 2 * y.pw;
  p.ypd = y.pd; p.yph = y.ph; p.ypw = y.pw; p.y_ld = y.ld; p.y_coff = y.coff;
  p.Cout = c->Cout; p.act = c->act;
  if (c->res != nullptr) {
    TSP_CHECK(c->pool.ptr == nullptr && !fused_oc && !(c->up.ptr != nullptr) && (c->res_ld | c->res_coff) % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(c->res) & 15) == 0,
              "slab: a residual excludes the fused pool / OutConv / up-sampling and must be 16-byte aligned per pixel chunk");
    p.res = reinterpret_cast<const __nv_bfloat16*>(c->res);
    static const bool res_wide_on = [] { const char* e = getenv("TEDSPAD_RES_WIDE"); return e == nullptr || e[0] != '0'; }();
    static const bool res_pf_on = [] { const char* e = getenv("TEDSPAD_RES_PREFETCH"); return e == nullptr || e[0] != '0'; }();
    p.res_prefetch = res_pf_on ? 1 : 0;
    p.res_wide = res_wide_on && ((c->res_ld | c->res_coff) & 15) == 0 && (reinterpret_cast<uintptr_t>(c->res) & 31) == 0;
    p.res_ld = c->res_ld; p.res_coff = c->res_coff;
  }
  if (c->pool.ptr != nullptr) {
    const tedspad_tensor& q = c->pool;
    if (check_tensor(q, "slab.pool", 8)) return 1;
    TSP_CHECK((c->kind == TEDSPAD_SLAB_3X3 || c->kind == TEDSPAD_SLAB_3X3_STREAM || c->kind == TEDSPAD_SLAB_3X3_PAIR ||
               c->kind == TEDSPAD_SLAB_3X3_STREAM_PAIR) &&
                  y.D == 1 && q.D == 1 && q.pd == 0 &&
                  q.N == y.N && q.C == c->Cout &&
                  q.H == y.H / 2 && q.W == y.W / 2,
              "slab: fused MaxPool2d(2) output [%d,%d,%d,%d] does not match", q.N, q.H, q.W, q.C);
    p.pool = reinterpret_cast<__nv_bfloat16*>(q.ptr);
    p.PH = q.H; p.PW = q.W; p.pHp = q.H + 2 * q.ph; p.pWp = q.W + 2 * q.pw; p.pph = q.ph; p.ppw = q.pw;
    p.p_ld = q.ld; p.p_coff = q.coff;
  }
  if (fused_oc) {
    TSP_CHECK(c->oc_b && (c->oc_planes || c->oc_clip.ptr), "slab: fused OutConv needs oc_b and oc_planes or oc_clip");
    TSP_CHECK(c->Cout <= 64 && c->Cout % 4 == 0 && y.D == 1, "slab: fused OutConv needs a 2-D layer with Cout <= 64");
    p.oc_w = c->oc_w; p.oc_b = c->oc_b;
    p.oc_planes = reinterpret_cast<__nv_bfloat16*>(c->oc_planes);
    p.oc_frames = c->oc_frames;
    p.dv_T = make_div(1);
    p.oc_T = 1;
    if (c->oc_clip.ptr != nullptr) {
      const tedspad_tensor& e = c->oc_clip;
      if (check_tensor(e, "slab.oc_clip", 1)) return 1;
      TSP_CHECK(c->oc_T >= 1 && x.N % c->oc_T == 0 && e.N == x.N / c->oc_T && e.D == c->oc_T && e.H == y.H && e.W == y.W &&
                    e.C >= 3,
                "slab: oc_clip [%d,%d,%d,%d,%d] does not match %d frames of T=%d", e.N, e.D, e.H, e.W, e.C, x.N, c->oc_T);
      p.oc_clip = reinterpret_cast<__nv_bfloat16*>(e.ptr);
      p.oc_T = c->oc_T;
      p.dv_T = make_div(c->oc_T);
      p.cDp = e.D + 2 * e.pd; p.cHp = e.H + 2 * e.ph; p.cWp = e.W + 2 * e.pw;
      p.cpd = e.pd; p.cph = e.ph; p.cpw = e.pw; p.c_ld = e.ld; p.c_coff = e.coff;
    }
  }

  if (s2d_clip) {
    const tedspad_tensor& e = c->oc_clip;
    if (check_tensor(e, "slab.oc_clip", 1)) return 1;
    TSP_CHECK(c->oc_T >= 1 && x.N % c->oc_T == 0 && e.N == x.N / c->oc_T && e.D == c->oc_T && e.H == 2 * y.H && e.W == 2 * y.W &&
                  e.C >= 3,
              "slab kx: clip [%d,%d,%d,%d,%d] does not match %d space-to-depth frames of T=%d", e.N, e.D, e.H, e.W, e.C, x.N, c->oc_T);
    p.oc_clip = reinterpret_cast<__nv_bfloat16*>(e.ptr);
    p.oc_T = c->oc_T;
    p.dv_T = make_div(c->oc_T);
    p.cDp = e.D + 2 * e.pd; p.cHp = e.H + 2 * e.ph; p.cWp = e.W + 2 * e.pw;
    p.cpd = e.pd; p.cph = e.ph; p.cpw = e.pw; p.c_ld = e.ld; p.c_coff = e.coff;
  }
  const bool has_up = c->up.ptr != nullptr;
  if (has_up) {
    const tedspad_tensor& u = c->up;
    if (check_tensor(u, "slab.up", 8)) return 1;
    p.up = reinterpret_cast<const __nv_bfloat16*>(u.ptr);
    p.up_cb_first = P.up_cb_first;
    p.up_H = u.H; p.up_W = u.W; p.up_Hp = u.H + 2 * u.ph; p.up_Wp = u.W + 2 * u.pw; p.up_ph = u.ph; p.up_pw = u.pw;
    p.up_ld = u.ld; p.up_coff = u.coff;
    // same geometry as tedspad_upsample2x: x2 bilinear, align_corners=True, centred with F.pad (unet_parts.py:50,57-63)
    p.up_UH = 2 * u.H; p.up_UW = 2 * u.W;
    p.up_offy = (x.H - p.up_UH) / 2; p.up_offx = (x.W - p.up_UW) / 2;
    p.up_sy = p.up_UH > 1 ? static_cast<float>(u.H - 1) / static_cast<float>(p.up_UH - 1) : 0.f;
    p.up_sx = p.up_UW > 1 ? static_cast<float>(u.W - 1) / static_cast<float>(p.up_UW - 1) : 0.f;
    p.slab_w = P.box[1]; p.slab_px = P.box[1] * P.box[2];
    p.slab_w_magic = static_cast<uint32_t>((0x100000000ULL + p.slab_w - 1) / p.slab_w);
  } else {
    p.up_cb_first = 1 << 20;
  }

  // staged TMA stores (see SlabKParams::tmY): plain epilogue, single CTA, 16 x 16 tiles of 64 output channels, tiles that
  // do not run across images; the staging area (8 warps x 2 x 4 KB) sits behind everything make_plan laid out
  int smem_bytes = P.smem_bytes;
  {
    static const bool tst_on = [] { const char* e = getenv("TEDSPAD_TMA_STORE"); return e == nullptr || e[0] != '0'; }();
    const int tst_off = static_cast<int>(round_up(P.smem_bytes - 1024, 1024));   // offset from the 1024-aligned base
    if (tst_on && !P.pair && !has_up && !fused_oc && c->res == nullptr && c->pool.ptr == nullptr && P.tm == 2 &&
        P.n_tile == 64 && c->Cout == 64 && P.num_n_tiles == 1 && P.stack_hp == 0 && y.ld % 8 == 0 && y.coff % 8 == 0 &&
        1024 + tst_off + 8 * 2 * 4096 <= SLAB_SMEM_BUDGET) {
      const uint64_t ydims[5] = {(uint64_t)c->Cout, (uint64_t)y.W, (uint64_t)y.H, (uint64_t)y.D, (uint64_t)y.N};
      const uint64_t px = (uint64_t)y.ld * 2;
      const uint64_t ystr[4] = {px, px * p.yWp, px * p.yWp * p.yHp, px * p.yWp * p.yHp * p.yDp};
      const uint32_t ybox[5] = {64, 8, 4, 1, 1};
      const uint8_t* ybase = reinterpret_cast<const uint8_t*>(y.ptr) +
                             2 * ((((int64_t)y.pd * p.yHp + y.ph) * p.yWp + y.pw) * y.ld + y.coff);
      // (a tensor the map cannot describe - base not 16-byte aligned, a stride beyond 2^40 - keeps the direct stores)
      if ((reinterpret_cast<uintptr_t>(ybase) & 15) == 0 && px * p.yWp * p.yHp * p.yDp < (1ULL << 40) &&
          encode_tmap_5d_bf16(&p.tmY, ybase, ydims, ystr, ybox, true) == 0) {
        p.tst = 1;
        p.tst_off = tst_off;
        smem_bytes = 1024 + tst_off + 8 * 2 * 4096;
      }
    }
  }

  if (device_once(ONCE_SLAB_ATTR)) {   // per device: the opt-in to > 48 KB of dynamic shared memory
    cudaError_t e = cudaFuncSetAttribute(conv_slab_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB_SMEM_BUDGET);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_slab_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB_SMEM_BUDGET);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_slab_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB_SMEM_BUDGET);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_slab_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB_SMEM_BUDGET);
    if (e != cudaSuccess) {
      device_once_reset(ONCE_SLAB_ATTR);
      set_error("cudaFuncSetAttribute(slab smem) failed: %s", cudaGetErrorString(e));
      return 2;
    }
  }
  int ctas = c->max_ctas > 0 ? c->max_ctas : num_sms();
  ctas = std::max(1, std::min(ctas, p.total_tiles));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (P.pair) {
    // clusters of two CTAs (one TPC); an even grid keeps the two tile loops of a pair in lock step
    ctas &= ~1;
    TSP_CHECK(ctas >= 2 && !has_up, "slab pair: needs at least two CTAs and no fused up-sampling");
    if (plan_e16(P, fused_oc))
      TSP_CUDA(launch_kernel(conv_slab_kernel<false, true, true>, dim3(ctas), dim3(SLAB_THREADS_E16), P.smem_bytes, st, p, 2));
    else
      TSP_CUDA(launch_kernel(conv_slab_kernel<false, true, false>, dim3(ctas), dim3(SLAB_THREADS), P.smem_bytes, st, p, 2));
  } else if (has_up) {
    TSP_CUDA(launch_kernel(conv_slab_kernel<true, false, false>, dim3(ctas), dim3(SLAB_THREADS_UP), P.smem_bytes, st, p));
  } else {
    TSP_CUDA(launch_kernel(conv_slab_kernel<false, false, false>, dim3(ctas), dim3(SLAB_THREADS), smem_bytes, st, p));
  }
  TSP_CUDA(cudaGetLastError());
  return 0;
}
