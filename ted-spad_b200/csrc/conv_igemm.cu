// Implicit-GEMM convolution on Blackwell tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
//   out[m, co] = act( sum_{tap, ci} in[pixel(m) + tap, ci] * w[co, tap, ci] + bias[co] (+ res[m, co]) )
//
// GEMM view: M = output pixels (128 per tile = the 128 TMEM lanes), N = output channels
// (n_tile <= 256 TMEM columns), K = taps x input channels in blocks of 64 bf16 (one 128-byte swizzle
// row per pixel).  One persistent CTA per SM walks tiles; its warps are specialised:
//
//   warp 0      TMA producer   weights (B) always; activations (A) in FLAT feed
//   warp 1      MMA issuer     one thread issues 4 x tcgen05.mma (K=16) per 64-wide K block
//   warp 2      TMEM allocator
//   warps 4-7   epilogue       tcgen05.ld -> +bias (+residual) -> ReLU/sigmoid -> bf16/fp32 stores
//   warps 8-11  A gatherers    GATHER feed only: cp.async 16-byte im2col into the swizzled layout
//
// FLAT feed (stride-1 "same" convolutions, the anonymizer UNet): activations live in a
// zero-haloed channels-last buffer, so for a tile of 128 consecutive *padded* pixel indices every
// filter tap is the same 2-D [rows, channels] slab shifted by a constant row offset: one TMA box
// per (tap, 64-channel block), no im2col arithmetic at all.  Halo pixels are computed like any
// other row and written back as zeros, which keeps the invariant for the next layer.
// GATHER feed (strided / asymmetric-pad / small-channel 3-D convolutions of the encoders): four
// producer warps build the same swizzled tile with predicated (zero-filling) cp.async.
//
// Pipelines: smem ring full[]/empty[] (producers <-> MMA) and a 2-deep TMEM accumulator ring
// tfull[]/tempty[] (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.h"
#include "ptx.cuh"

namespace tsp {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_TAPS = 27;
constexpr int GATHER_LAG = 2;  // cp.async groups kept in flight per gather thread
constexpr int SMEM_BUDGET = 227 * 1024;

struct ConvKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int feed;  // TEDSPAD_FEED_FLAT_TMA or TEDSPAD_FEED_GATHER
  int M_total, num_m_tiles, num_n_tiles, n_tile, num_kb, stages, tmem_cols;
  int Cout;
  // FLAT
  int cin_chunks;
  int tap_off[MAX_TAPS];
  // input geometry (GATHER) / padded geometry (FLAT border mask uses the y* fields)
  const __nv_bfloat16* x;
  int xD, xH, xW, xDp, xHp, xWp, xpd, xph, xpw, x_ld, x_coff, cin8, ntaps;
  int kh, kw, sd, sh, sw, fpd, fph, fpw;
  // output
  void* y;
  int OD, OH, OW, yDp, yHp, yWp, ypd, yph, ypw, y_ld, y_coff, y_fp32;
  const float* bias;
  const __nv_bfloat16* res;
  int res_ld, res_coff, act;
  // output column segments (same pixel geometry, different buffers / channel slices): segment s holds columns
  // [seg_begin[s], seg_begin[s] + seg_cout[s]); a single-destination convolution is one segment starting at 0
  int nseg;
  int seg_begin[3], seg_cout[3], seg_ld[3], seg_coff[3];
  void* seg_y[3];
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == TEDSPAD_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TEDSPAD_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

__global__ void __launch_bounds__(384, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int S = p.stages;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.n_tile) * 128u;
  uint8_t* smA = smem;
  uint8_t* smB = smem + S * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smB + S * b_stage_bytes);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool gather = p.feed == TEDSPAD_FEED_GATHER;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    if (!gather) tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, gather ? 1 + 4 : 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();                // (this kernel's prologue above touched no tensor memory)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        const int m0 = mt * BLOCK_M, n0 = nt * p.n_tile;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          if (!gather) {
            mbar_arrive_expect_tx(full + s, A_STAGE_BYTES + b_stage_bytes);
            const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            tma_load_2d(smA + s * A_STAGE_BYTES, &p.tmA, full + s, cc * BLOCK_K, m0 + p.tap_off[tap]);
          } else {
            mbar_arrive_expect_tx(full + s, b_stage_bytes);
          }
          tma_load_2d(smB + s * b_stage_bytes, &p.tmB, full + s, kb * BLOCK_K, n0);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, p.n_tile);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty + as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.n_tile);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smB + s * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty + s);  // frees the smem slot once these MMAs have read it
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit(tfull + as);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    const int row = ew * 32 + lane;
    bool seg_vec[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) seg_vec[i] = ((p.seg_ld[i] | p.seg_coff[i]) & 7) == 0;
    const bool rvec_ok = p.res != nullptr && ((p.res_ld | p.res_coff) & 7) == 0;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int m = mt * BLOCK_M + row, n0 = nt * p.n_tile;
      const bool row_ok = m < p.M_total;
      bool interior = true;
      long long pix;
      if (!gather) {
        int t = m;
        const int wq = t % p.yWp; t /= p.yWp;
        const int hq = t % p.yHp; t /= p.yHp;
        const int dq = t % p.yDp;
        interior = (wq >= p.ypw) && (wq < p.ypw + p.OW) && (hq >= p.yph) && (hq < p.yph + p.OH) &&
                   (dq >= p.ypd) && (dq < p.ypd + p.OD);
        pix = m;
      } else {
        int t = m;
        const int ow = t % p.OW; t /= p.OW;
        const int oh = t % p.OH; t /= p.OH;
        const int od = t % p.OD;
        const int n = t / p.OD;
        pix = ((static_cast<long long>(n) * p.yDp + od + p.ypd) * p.yHp + oh + p.yph) * p.yWp + ow + p.ypw;
      }
      const long long roff = pix * p.res_ld + p.res_coff;

      mbar_wait(tfull + as, aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as * p.n_tile);
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + c, v);
        tmem_ld_wait();
        const int cg = n0 + c;
        // destination of this 16-column chunk (segments start at multiples of 16)
        int sg = 0;
        if (p.nseg > 1 && cg >= p.seg_begin[1]) sg = (p.nseg > 2 && cg >= p.seg_begin[2]) ? 2 : 1;
        const int cl = cg - p.seg_begin[sg];          // column inside the segment
        const int s_cout = p.seg_cout[sg];
        const bool vec_ok = seg_vec[sg];
        const long long yoff = pix * p.seg_ld[sg] + p.seg_coff[sg];
        if (row_ok && cl < s_cout) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + __ldg(p.bias + cg + i);
          if (p.res != nullptr) {
            if (rvec_ok && cl + 16 <= s_cout) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + roff + cg);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 r = __ldg(rp + h);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 q = __bfloat1622float2(r2[i]);
                  f[h * 8 + 2 * i] += q.x;
                  f[h * 8 + 2 * i + 1] += q.y;
                }
              }
            } else {
              for (int i = 0; i < 16; ++i)
                if (cl + i < s_cout) f[i] += __bfloat162float(p.res[roff + cg + i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = interior ? apply_act(f[i], p.act) : 0.f;
          if (p.y_fp32) {
            float* yp = reinterpret_cast<float*>(p.seg_y[sg]) + yoff + cl;
            for (int i = 0; i < 16; ++i)
              if (cl + i < s_cout) yp[i] = f[i];
          } else {
            __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.seg_y[sg]) + yoff + cl;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (vec_ok && cl + h * 8 + 8 <= s_cout) {
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[h * 8 + 2 * i], f[h * 8 + 2 * i + 1]);
                *reinterpret_cast<uint4*>(yp + h * 8) = o;
              } else {
                for (int i = 0; i < 8; ++i)
                  if (cl + h * 8 + i < s_cout) yp[h * 8 + i] = __float2bfloat16_rn(f[h * 8 + i]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + as);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else if (warp >= 8 && gather) {
    // ------------------------------------------------------------- A gatherers
    // lane -> (16-byte chunk j of the 64-wide K block, 4 rows per warp per step); 8 steps x 16 rows.
    const int gw = warp - 8;
    const int j = lane & 7;
    const int rsub = gw * 4 + (lane >> 3);
    const long long x_pix_stride = p.x_ld;
    int s = 0;
    uint32_t ph = 0;
    int issued = 0;  // K blocks committed so far (whole kernel)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles;
      const int m0 = mt * BLOCK_M;
      int nb[8], id0[8], ih0[8], iw0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int t = m0 + i * 16 + rsub;
        const bool ok = t < p.M_total;
        const int ow = t % p.OW; t /= p.OW;
        const int oh = t % p.OH; t /= p.OH;
        const int od = t % p.OD;
        const int n = t / p.OD;
        nb[i] = ok ? n : -1;
        id0[i] = od * p.sd - p.fpd;
        ih0[i] = oh * p.sh - p.fph;
        iw0[i] = ow * p.sw - p.fpw;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(empty + s, ph ^ 1);
        const int kc = kb * 8 + j;
        const int tap = kc / p.cin8;
        const int c8 = kc - tap * p.cin8;
        const bool k_ok = tap < p.ntaps;
        const int tkw = tap % p.kw;
        const int t2 = tap / p.kw;
        const int tkh = t2 % p.kh;
        const int tkd = t2 / p.kh;
        const uint32_t a_stage = smem_u32(smA + s * A_STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int id = id0[i] + tkd, ih = ih0[i] + tkh, iw = iw0[i] + tkw;
          const bool ok = k_ok && nb[i] >= 0 && static_cast<unsigned>(id) < static_cast<unsigned>(p.xD) &&
                          static_cast<unsigned>(ih) < static_cast<unsigned>(p.xH) &&
                          static_cast<unsigned>(iw) < static_cast<unsigned>(p.xW);
          const long long pix =
              ((static_cast<long long>(nb[i]) * p.xDp + id + p.xpd) * p.xHp + ih + p.xph) * p.xWp + iw + p.xpw;
          const __nv_bfloat16* src = ok ? (p.x + pix * x_pix_stride + p.x_coff + c8 * 8) : p.x;
          const int r = i * 16 + rsub;
          const uint32_t dst = a_stage + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
          cp_async_16_zfill(dst, src, ok ? 16u : 0u);
        }
        cp_async_commit();
        ++issued;
        if (issued > GATHER_LAG) {
          cp_async_wait<GATHER_LAG>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full + ((issued - 1 - GATHER_LAG) % S));
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const int first = issued > GATHER_LAG ? issued - GATHER_LAG : 0;
      for (int q = first; q < issued; ++q) mbar_arrive(full + (q % S));
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

static int pick_n_tile(int cout_pad) {
  const int nt = (cout_pad + 255) / 256;
  const int n = static_cast<int>(round_up((cout_pad + nt - 1) / nt, 16));
  return n;
}


}  // namespace tsp

using namespace tsp;

extern "C" int tedspad_conv_forward(const tedspad_conv* c, void* stream_v) {
  TSP_CHECK(c != nullptr, "conv: null descriptor");
  const tedspad_tensor& x = c->x;
  const tedspad_tensor& y = c->y;
  if (check_tensor(x, "conv.x", 8) || check_tensor(y, "conv.y", 1)) return 1;
  TSP_CHECK(c->w && c->bias, "conv: null weights/bias");
  TSP_CHECK(c->kd >= 1 && c->kh >= 1 && c->kw >= 1 && c->sd >= 1 && c->sh >= 1 && c->sw >= 1, "conv: bad kernel/stride");
  TSP_CHECK(x.N == y.N, "conv: batch mismatch %d vs %d", x.N, y.N);
  const tedspad_tensor* extra[2] = {&c->y2, &c->y3};
  const int extra_begin[2] = {c->y2_begin, c->y3_begin};
  int nseg = 1;
  for (int i = 0; i < 2; ++i) {
    if (extra[i]->ptr == nullptr) break;
    const tedspad_tensor& e = *extra[i];
    if (check_tensor(e, "conv.y2/y3", 1)) return 1;
    const int prev_end = i == 0 ? y.C : extra_begin[0] + extra[0]->C;
    TSP_CHECK(nseg == i + 1 && extra_begin[i] % 16 == 0 && extra_begin[i] >= prev_end && c->res == nullptr && !c->y_fp32 &&
                  e.N == y.N && e.D == y.D && e.H == y.H && e.W == y.W && e.pd == y.pd && e.ph == y.ph && e.pw == y.pw &&
                  extra_begin[i] + e.C <= c->Cout,
              "conv: extra destination %d (columns %d..%d of %d) must follow the previous one at a multiple of 16 and "
              "share y's pixel geometry", i + 2, extra_begin[i], extra_begin[i] + e.C, c->Cout);
    ++nseg;
  }
  TSP_CHECK(nseg > 1 ? y.C <= c->Cout : y.C == c->Cout, "conv: y.C %d != Cout %d", y.C, c->Cout);
  TSP_CHECK(x.C % 8 == 0, "conv: x.C=%d must be a multiple of 8 (pad input channels)", x.C);
  const int ntaps = c->kd * c->kh * c->kw;
  TSP_CHECK(c->K_pad % BLOCK_K == 0 && c->K_pad >= ntaps * x.C, "conv: K_pad=%d invalid for %d taps x %d ch", c->K_pad,
            ntaps, x.C);
  TSP_CHECK(c->Cout_pad % 16 == 0 && c->Cout_pad >= c->Cout, "conv: Cout_pad=%d invalid", c->Cout_pad);
  // sanity of the output extents: the last window must start inside the (front-padded) input
  auto fits = [](int in, int k, int s, int pf, int out) { return out >= 1 && pf >= 0 && pf < k && (out - 1) * s - pf < in; };
  TSP_CHECK(fits(x.D, c->kd, c->sd, c->pd, y.D) && fits(x.H, c->kh, c->sh, c->ph, y.H) &&
                fits(x.W, c->kw, c->sw, c->pw, y.W),
            "conv: output extents (%d,%d,%d) inconsistent with input (%d,%d,%d)", y.D, y.H, y.W, x.D, x.H, x.W);
  const bool flat_legal = c->sd == 1 && c->sh == 1 && c->sw == 1 && (c->kd & 1) && (c->kh & 1) && (c->kw & 1) &&
                          c->pd == c->kd / 2 && c->ph == c->kh / 2 && c->pw == c->kw / 2 && x.C % 64 == 0 &&
                          x.D == y.D && x.H == y.H && x.W == y.W && x.pd == y.pd && x.ph == y.ph && x.pw == y.pw &&
                          x.pd >= c->pd && x.ph >= c->ph && x.pw >= c->pw && ntaps <= MAX_TAPS &&
                          c->K_pad == ntaps * x.C;
  int feed = c->feed;
  if (feed == TEDSPAD_FEED_AUTO) feed = flat_legal ? TEDSPAD_FEED_FLAT_TMA : TEDSPAD_FEED_GATHER;
  TSP_CHECK(feed == TEDSPAD_FEED_GATHER || (feed == TEDSPAD_FEED_FLAT_TMA && flat_legal),
            "conv: FLAT feed requested but not legal for this layer");

  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.feed = feed;
  p.n_tile = c->n_tile > 0 ? c->n_tile : pick_n_tile(c->Cout_pad);
  TSP_CHECK(p.n_tile % 16 == 0 && p.n_tile >= 16 && p.n_tile <= 256 && c->Cout_pad % p.n_tile == 0,
            "conv: n_tile=%d does not tile Cout_pad=%d", p.n_tile, c->Cout_pad);
  p.num_n_tiles = c->Cout_pad / p.n_tile;
  p.num_kb = c->K_pad / BLOCK_K;
  p.Cout = c->Cout;
  const int64_t ypix = tensor_pixels(y);
  const int64_t m_total = feed == TEDSPAD_FEED_FLAT_TMA ? ypix : (int64_t)y.N * y.D * y.H * y.W;
  TSP_CHECK(m_total > 0 && m_total < (int64_t(1) << 31) - 256, "conv: M=%lld out of range", (long long)m_total);
  TSP_CHECK(tensor_pixels(x) * x.ld < (int64_t(1) << 40), "conv: input too large");
  p.M_total = (int)m_total;
  p.num_m_tiles = (int)((m_total + BLOCK_M - 1) / BLOCK_M);
  TSP_CHECK((int64_t)p.num_m_tiles * p.num_n_tiles < (int64_t(1) << 31), "conv: too many tiles");

  const int stage_bytes = A_STAGE_BYTES + p.n_tile * 128;
  p.stages = std::min(8, (SMEM_BUDGET - 2048) / stage_bytes);
  TSP_CHECK(p.stages >= GATHER_LAG + 1, "conv: not enough smem stages");
  int tc = 32;
  while (tc < 2 * p.n_tile) tc <<= 1;
  p.tmem_cols = tc;
  const int smem_bytes = p.stages * stage_bytes + 1024 + 256;

  p.x = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
  p.xD = x.D; p.xH = x.H; p.xW = x.W;
  p.xDp = x.D + 2 * x.pd; p.xHp = x.H + 2 * x.ph; p.xWp = x.W + 2 * x.pw;
  p.xpd = x.pd; p.xph = x.ph; p.xpw = x.pw;
  p.x_ld = x.ld; p.x_coff = x.coff; p.cin8 = x.C / 8; p.ntaps = ntaps;
  p.kh = c->kh; p.kw = c->kw; p.sd = c->sd; p.sh = c->sh; p.sw = c->sw;
  p.fpd = c->pd; p.fph = c->ph; p.fpw = c->pw;
  p.y = y.ptr;
  p.OD = y.D; p.OH = y.H; p.OW = y.W;
  p.yDp = y.D + 2 * y.pd; p.yHp = y.H + 2 * y.ph; p.yWp = y.W + 2 * y.pw;
  p.ypd = y.pd; p.yph = y.ph; p.ypw = y.pw;
  p.y_ld = y.ld; p.y_coff = y.coff; p.y_fp32 = c->y_fp32;
  p.nseg = nseg;
  p.seg_begin[0] = 0; p.seg_cout[0] = y.C; p.seg_ld[0] = y.ld; p.seg_coff[0] = y.coff; p.seg_y[0] = y.ptr;
  for (int i = 1; i < 3; ++i) {
    const bool on = i < nseg;
    const tedspad_tensor& e = on ? *extra[i - 1] : y;
    p.seg_begin[i] = on ? extra_begin[i - 1] : (1 << 30);
    p.seg_cout[i] = on ? e.C : 0; p.seg_ld[i] = e.ld; p.seg_coff[i] = e.coff; p.seg_y[i] = e.ptr;
  }
  p.bias = c->bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(c->res);
  p.res_ld = c->res_ld; p.res_coff = c->res_coff;
  p.act = c->act;

  if (encode_tmap_2d_bf16(&p.tmB, c->w, (uint64_t)c->K_pad, (uint64_t)c->Cout_pad, (uint64_t)c->K_pad * 2, BLOCK_K,
                          (uint32_t)p.n_tile))
    return 3;
  if (feed == TEDSPAD_FEED_FLAT_TMA) {
    p.cin_chunks = x.C / 64;
    int t = 0;
    for (int a = 0; a < c->kd; ++a)
      for (int b = 0; b < c->kh; ++b)
        for (int d = 0; d < c->kw; ++d) p.tap_off[t++] = ((a - c->pd) * p.xHp + (b - c->ph)) * p.xWp + (d - c->pw);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(x.ptr) + (size_t)x.coff * 2;
    if (encode_tmap_2d_bf16(&p.tmA, base, (uint64_t)x.C, (uint64_t)m_total, (uint64_t)x.ld * 2, BLOCK_K, BLOCK_M))
      return 3;
  }

  if (device_once(ONCE_IGEMM_ATTR)) {   // per device: the opt-in to > 48 KB of dynamic shared memory
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
    if (e != cudaSuccess) {
      device_once_reset(ONCE_IGEMM_ATTR);
      set_error("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return 2;
    }
  }

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  int ctas = c->max_ctas > 0 ? c->max_ctas : num_sms();
  ctas = std::max(1, std::min(ctas, total_tiles));
  const int threads = feed == TEDSPAD_FEED_GATHER ? 384 : 256;
  TSP_CUDA(launch_kernel(conv_igemm_kernel, dim3(ctas), dim3(threads), smem_bytes, reinterpret_cast<cudaStream_t>(stream_v), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}
