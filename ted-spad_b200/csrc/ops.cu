// HBM-bound side kernels of the hot path: pooling, up-sampling into concat slices, the OutConv +
// sigmoid + raw-reshape scatter, feature mean-pooling, uint8 preprocessing and layout adapters.
// All activations are channels-last bf16; every kernel moves 16 bytes (8 channels) per access.
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace tsp {

struct TView {  // device-side copy of tedspad_tensor with derived padded extents
  uint8_t* ptr;
  int N, D, H, W, C, pd, ph, pw, ld, coff, Dp, Hp, Wp;
};

static TView make_view(const tedspad_tensor& t) {
  TView v;
  v.ptr = reinterpret_cast<uint8_t*>(t.ptr);
  v.N = t.N; v.D = t.D; v.H = t.H; v.W = t.W; v.C = t.C;
  v.pd = t.pd; v.ph = t.ph; v.pw = t.pw; v.ld = t.ld; v.coff = t.coff;
  v.Dp = t.D + 2 * t.pd; v.Hp = t.H + 2 * t.ph; v.Wp = t.W + 2 * t.pw;
  return v;
}

__device__ __forceinline__ long long pix_index(const TView& v, int n, int d, int h, int w) {
  return ((static_cast<long long>(n) * v.Dp + d + v.pd) * v.Hp + h + v.ph) * v.Wp + w + v.pw;
}
__device__ __forceinline__ const __nv_bfloat16* elem_ptr(const TView& v, long long pix, int c) {
  return reinterpret_cast<const __nv_bfloat16*>(v.ptr) + pix * v.ld + v.coff + c;
}
__device__ __forceinline__ __nv_bfloat16* elem_ptr_w(const TView& v, long long pix, int c) {
  return reinterpret_cast<__nv_bfloat16*>(v.ptr) + pix * v.ld + v.coff + c;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

constexpr long long ROWS_PER_BLOCK = 1;  // >1 was measured: no L1-reuse gain, less parallelism on small maps

// bf16 pair -> two fp32 (a bf16 is the upper half of an fp32: one shift / one mask)
__device__ __forceinline__ void bf2_to_f32(uint32_t v, float& lo, float& hi) {
  lo = __uint_as_float(v << 16);
  hi = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ uint32_t hmax2_bf16(uint32_t a, uint32_t b) {
  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}
// it / d for it < 2^16 with magic = ceil(2^32 / d) (exact in that range for d <= 2^16)
__device__ __forceinline__ int fast_div(int it, uint32_t magic) { return static_cast<int>(__umulhi(static_cast<uint32_t>(it), magic)); }

// ------------------------------------------------------------------ max pool
struct PoolP {
  TView x, y;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw, zero_pad;
  uint32_t c8_magic;
  long long total;  // output rows N*OD*OH
};

// One block per output row (n, od, oh): the row decode is done once per block, the threads walk the
// (ow, 8-channel group) items of the row with 32-bit arithmetic only.  KD/KH/KW > 0: compile-time window, so
// the whole window's loads are issued before the first max (the runtime-window loop serialises on load latency:
// 0.26 ms for the 27-tap Inception pool of Mixed_3b vs a 0.03 ms DRAM floor); 0 = runtime window (p.kd/kh/kw).
__device__ __forceinline__ void max8(uint4& m, const uint4& q) {
  m.x = hmax2_bf16(m.x, q.x); m.y = hmax2_bf16(m.y, q.y);
  m.z = hmax2_bf16(m.z, q.z); m.w = hmax2_bf16(m.w, q.w);
}

template <int KD, int KH, int KW>
__global__ void __launch_bounds__(256) maxpool_kernel(const PoolP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int kd = KD > 0 ? KD : p.kd, kh = KH > 0 ? KH : p.kh, kw = KW > 0 ? KW : p.kw;
  const int c8n = p.y.C >> 3;
  const int items = p.y.W * c8n;
  const uint4 neg = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // -inf pairs
  for (long long row = blockIdx.x; row < p.total; row += gridDim.x) {
    int t = static_cast<int>(row);
    const int oh = t % p.y.H; t /= p.y.H;
    const int od = t % p.y.D;
    const int n = t / p.y.D;
    const int id0 = od * p.sd - p.pd, ih0 = oh * p.sh - p.ph;
    const __nv_bfloat16* xn = reinterpret_cast<const __nv_bfloat16*>(p.x.ptr) + p.x.coff;
    __nv_bfloat16* yrow = elem_ptr_w(p.y, pix_index(p.y, n, od, oh, 0), 0);
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int ow = fast_div(it, p.c8_magic), c8 = it - ow * c8n;
      const int iw0 = ow * p.sw - p.pw;
      // max on packed bf16 pairs (exact: max commutes with rounding)
      uint4 m = neg;
      bool any_oob = false;
#pragma unroll
      for (int a = 0; a < kd; ++a) {
        const int id = id0 + a;
        const bool okd = static_cast<unsigned>(id) < static_cast<unsigned>(p.x.D);
#pragma unroll
        for (int b = 0; b < kh; ++b) {
          const int ih = ih0 + b;
          const bool okh = okd && static_cast<unsigned>(ih) < static_cast<unsigned>(p.x.H);
          const long long rowpix = okh ? pix_index(p.x, n, id, ih, 0) : 0;
#pragma unroll
          for (int c = 0; c < kw; ++c) {
            const int iw = iw0 + c;
            const bool ok = okh && static_cast<unsigned>(iw) < static_cast<unsigned>(p.x.W);
            uint4 q = neg;
            if (ok) q = __ldg(reinterpret_cast<const uint4*>(xn + (rowpix + iw) * p.x.ld + c8 * 8));
            max8(m, q);
            any_oob |= !ok;
          }
        }
      }
      if (any_oob && p.zero_pad) max8(m, make_uint4(0u, 0u, 0u, 0u));
      *reinterpret_cast<uint4*>(yrow + static_cast<long long>(ow) * p.y.ld + c8 * 8) = m;
    }
  }
}

// (3,3,3) stride-1 pad-1 pooling (the Inception pool branch, i3d.py:138-139 through MaxPool3dSamePadding): one
// thread per (n, od, oh, 8-channel group) walks its output row with a sliding window of three column maxima
// (column = the 3x3 (d,h) taps at one w), i.e. 9 loads per output instead of 27.  The (d,h) bounds are per-thread
// constants; consecutive threads own consecutive channel groups, so every load instruction is coalesced.
__global__ void __launch_bounds__(256) maxpool333_s1_kernel(const PoolP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int c8n = p.y.C >> 3;
  const long long total = p.total * c8n;
  const uint4 neg = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const uint4 oob_col = p.zero_pad ? zero : neg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % c8n);
    int t = static_cast<int>(idx / c8n);
    const int oh = t % p.y.H; t /= p.y.H;
    const int od = t % p.y.D;
    const int n = t / p.y.D;
    const __nv_bfloat16* xn = reinterpret_cast<const __nv_bfloat16*>(p.x.ptr) + p.x.coff + c8 * 8;
    const __nv_bfloat16* rp[9];
    bool dh_oob = false;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int id = od - 1 + a, ih = oh - 1 + b;
        const bool ok = static_cast<unsigned>(id) < static_cast<unsigned>(p.x.D) &&
                        static_cast<unsigned>(ih) < static_cast<unsigned>(p.x.H);
        rp[a * 3 + b] = ok ? xn + pix_index(p.x, n, id, ih, 0) * p.x.ld : nullptr;
        dh_oob |= !ok;
      }
    const uint4 col_floor = (dh_oob && p.zero_pad) ? zero : neg;
    auto column = [&](int iw) {
      uint4 m = col_floor;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        uint4 q = neg;
        if (rp[i] != nullptr) q = __ldg(reinterpret_cast<const uint4*>(rp[i] + static_cast<long long>(iw) * p.x.ld));
        max8(m, q);
      }
      return m;
    };
    __nv_bfloat16* yrow = elem_ptr_w(p.y, pix_index(p.y, n, od, oh, 0), c8 * 8);
    uint4 c0 = oob_col, c1 = column(0);
#pragma unroll 2
    for (int ow = 0; ow < p.y.W; ++ow) {
      const uint4 c2 = ow + 1 < p.x.W ? column(ow + 1) : oob_col;
      uint4 m = c0;
      max8(m, c1);
      max8(m, c2);
      *reinterpret_cast<uint4*>(yrow + static_cast<long long>(ow) * p.y.ld) = m;
      c0 = c1;
      c1 = c2;
    }
  }
}

// Register-blocked form of the same pooling: a thread owns a 2 x 2 block of (od, oh) output rows and walks them along
// w together.  Per column it loads the 4 x 4 (d, h) input rows the block's windows cover once - 4 loads per output
// instead of 9 - and reduces them separably (3 along h, then 3 along d).  The one-row kernel above is bound by its
// load requests (splitting its rows over MORE threads was measured slower: 0.031 -> 0.042 ms on the 14 x 14 maps).
__global__ void __launch_bounds__(256) maxpool333_s1_rb_kernel(const PoolP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int c8n = p.y.C >> 3;
  const int D2 = (p.y.D + 1) >> 1, H2 = (p.y.H + 1) >> 1;
  const long long total = static_cast<long long>(p.y.N) * D2 * H2 * c8n;
  const uint4 neg = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const uint4 oobv = p.zero_pad ? zero : neg;   // an out-of-range tap is a zero (MaxPool3dSamePadding) or absent
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % c8n);
    int t = static_cast<int>(idx / c8n);
    const int oh0 = (t % H2) * 2; t /= H2;
    const int od0 = (t % D2) * 2;
    const int n = t / D2;
    const __nv_bfloat16* xn = reinterpret_cast<const __nv_bfloat16*>(p.x.ptr) + p.x.coff + c8 * 8;
    const __nv_bfloat16* rp[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int id = od0 - 1 + a, ih = oh0 - 1 + b;
        const bool ok = static_cast<unsigned>(id) < static_cast<unsigned>(p.x.D) &&
                        static_cast<unsigned>(ih) < static_cast<unsigned>(p.x.H);
        rp[a * 4 + b] = ok ? xn + pix_index(p.x, n, id, ih, 0) * p.x.ld : nullptr;
      }
    // column maxima of the four outputs at input column iw
    auto column = [&](int iw, uint4 (&col)[4]) {
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        v[i] = rp[i] != nullptr ? __ldg(reinterpret_cast<const uint4*>(rp[i] + static_cast<long long>(iw) * p.x.ld)) : oobv;
      uint4 mh[8];   // [a][j]: max over h of rows j .. j+2 at depth a
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 m = v[a * 4 + j];
          max8(m, v[a * 4 + j + 1]);
          max8(m, v[a * 4 + j + 2]);
          mh[a * 2 + j] = m;
        }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 m = mh[i * 2 + j];
          max8(m, mh[(i + 1) * 2 + j]);
          max8(m, mh[(i + 2) * 2 + j]);
          col[i * 2 + j] = m;
        }
    };
    __nv_bfloat16* yrow[4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        yrow[i * 2 + j] = (od0 + i < p.y.D && oh0 + j < p.y.H) ? elem_ptr_w(p.y, pix_index(p.y, n, od0 + i, oh0 + j, 0), c8 * 8)
                                                                 : nullptr;
    uint4 c0[4], c1[4], c2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c0[k] = oobv;
    column(0, c1);
    for (int ow = 0; ow < p.y.W; ++ow) {
      if (ow + 1 < p.x.W) {
        column(ow + 1, c2);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) c2[k] = oobv;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 m = c0[k];
        max8(m, c1[k]);
        max8(m, c2[k]);
        if (yrow[k] != nullptr) *reinterpret_cast<uint4*>(yrow[k] + static_cast<long long>(ow) * p.y.ld) = m;
        c0[k] = c1[k];
        c1[k] = c2[k];
      }
    }
  }
}

// ------------------------------------------------- bilinear x2, align_corners
struct UpP {
  TView x, y;
  int UH, UW, offy, offx;  // up-sampled size and F.pad offsets inside y
  float sy, sx;
  uint32_t c8_magic;
  long long total;  // N * bands of UP_ROWS output rows
};

// One block per UP_ROWS consecutive output rows of one image; a thread owns one (ow, 8-channel group) column of
// that band and walks down it.  The interpolation is separable - out = ly0*(lx0*a + lx1*b) + ly1*(lx0*c + lx1*d),
// the form aten's upsample_bilinear2d uses - so the horizontally interpolated source rows are kept in registers
// and re-used by the 2-3 output rows between them: ~1.25 16-byte loads per output instead of 4 and half the ALU
// work of the four-tap form (the kernel was instruction-bound at 2/3 of the HBM write ceiling).
constexpr int UP_ROWS = 8;

__device__ __forceinline__ void up_hrow(const __nv_bfloat16* r, long long o0, long long o1, float lx0, float lx1, float (&h)[8]) {
  const uint4 qa = __ldg(reinterpret_cast<const uint4*>(r + o0));
  const uint4 qb = __ldg(reinterpret_cast<const uint4*>(r + o1));
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&qa);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&qb);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a0, a1, b0, b1;
    bf2_to_f32(pa[i], a0, a1); bf2_to_f32(pb[i], b0, b1);
    h[2 * i] = fmaf(lx1, b0, lx0 * a0);
    h[2 * i + 1] = fmaf(lx1, b1, lx0 * a1);
  }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const UpP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int c8n = p.y.C >> 3;
  const int items = p.y.W * c8n;
  const int bands = (p.y.H + UP_ROWS - 1) / UP_ROWS;
  for (long long bi = blockIdx.x; bi < p.total; bi += gridDim.x) {
    const int n = static_cast<int>(bi / bands);
    const int oh0 = static_cast<int>(bi - static_cast<long long>(n) * bands) * UP_ROWS;
    const int oh1 = min(oh0 + UP_ROWS, p.y.H);
    const __nv_bfloat16* xn = elem_ptr(p.x, pix_index(p.x, n, 0, 0, 0), 0);
    const long long x_row = static_cast<long long>(p.x.Wp) * p.x.ld;
    __nv_bfloat16* yn = elem_ptr_w(p.y, pix_index(p.y, n, 0, oh0, 0), 0);
    const long long y_row = static_cast<long long>(p.y.Wp) * p.y.ld;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int ow = fast_div(it, p.c8_magic), c8 = it - ow * c8n;
      const int ux = ow - p.offx;
      const bool col_in = ux >= 0 && ux < p.UW;
      const float fx = p.sx * ux;
      const int x0 = col_in ? static_cast<int>(fx) : 0;
      const int x1 = x0 + (x0 < p.x.W - 1 ? 1 : 0);
      const float lx1 = fx - x0, lx0 = 1.f - lx1;
      const long long o0 = static_cast<long long>(x0) * p.x.ld + c8 * 8, o1 = static_cast<long long>(x1) * p.x.ld + c8 * 8;
      __nv_bfloat16* yp = yn + static_cast<long long>(ow) * p.y.ld + c8 * 8;
      float h0[8], h1[8];
      int cur = -2;   // source row held in h0 (h1 holds row min(cur + 1, H - 1))
      for (int oh = oh0; oh < oh1; ++oh, yp += y_row) {
        const int uy = oh - p.offy;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (col_in && uy >= 0 && uy < p.UH) {
          const float fy = p.sy * uy;
          const int y0 = static_cast<int>(fy);
          const int y1 = y0 + (y0 < p.x.H - 1 ? 1 : 0);
          const float ly1 = fy - y0, ly0 = 1.f - ly1;
          if (y0 != cur) {
            if (y0 == cur + 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) h0[i] = h1[i];
            } else {
              up_hrow(xn + y0 * x_row, o0, o1, lx0, lx1, h0);
            }
            if (y1 != y0) {
              up_hrow(xn + y1 * x_row, o0, o1, lx0, lx1, h1);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) h1[i] = h0[i];
            }
            cur = y0;
          }
          uint32_t* po = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            po[i] = cvt_bf16x2(fmaf(ly1, h1[2 * i], ly0 * h0[2 * i]), fmaf(ly1, h1[2 * i + 1], ly0 * h0[2 * i + 1]), false);
        }
        *reinterpret_cast<uint4*>(yp) = out;
      }
    }
  }
}

// --------------------------------- OutConv 1x1 (C->3) + sigmoid + plane scatter
struct OutP {
  TView x, y;
  const float* w;
  const float* b;
  float* frames;
  int T;
  long long total;  // frames*H*W pixels
};

// 8 lanes per pixel: each lane owns 16-byte channel chunks lane, lane+8, ...; partial dot products
// are combined with warp shuffles.
__global__ void __launch_bounds__(256) outconv_sigmoid_kernel(const OutP p) {
  extern __shared__ float sw[];  // [3][C]
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 3 * p.x.C; i += blockDim.x) sw[i] = p.w[i];   // constants: before the wait
  __syncthreads();
  pdl_wait();
  const int sub = threadIdx.x & 7;
  const int chunks = p.x.C >> 3;
  const long long gstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 3);
  const long long iters = (p.total + gstride - 1) / gstride;
  long long pixi = blockIdx.x * static_cast<long long>(blockDim.x >> 3) + (threadIdx.x >> 3);
  for (long long it = 0; it < iters; ++it, pixi += gstride) {
    const bool ok = pixi < p.total;
    float acc[3] = {0.f, 0.f, 0.f};
    int w_ = 0, h_ = 0, fr = 0;
    if (ok) {
      long long t = pixi;
      w_ = static_cast<int>(t % p.x.W); t /= p.x.W;
      h_ = static_cast<int>(t % p.x.H);
      fr = static_cast<int>(t / p.x.H);
      const long long pin = pix_index(p.x, fr, 0, h_, w_);
      for (int ch = sub; ch < chunks; ch += 8) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(elem_ptr(p.x, pin, ch * 8))), f);
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float* wr = sw + o * p.x.C + ch * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[o] = fmaf(f[i], wr[i], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 4);
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
    }
    if (ok && sub < 3) {
      const float v = 1.f / (1.f + __expf(-(acc[sub] + p.b[sub])));
      const int bclip = fr / p.T, tf = fr - bclip * p.T;
      const int plane = 3 * tf + sub;            // dali_extraction.py:173 raw reshape
      const int ce = plane / p.T, te = plane - ce * p.T;
      *elem_ptr_w(p.y, pix_index(p.y, bclip, te, h_, w_), ce) = __float2bfloat16_rn(v);
      if (p.frames) p.frames[((static_cast<long long>(fr) * 3 + sub) * p.x.H + h_) * p.x.W + w_] = v;
    }
  }
}

// ------------------------------------------------------------ feature mean pool
struct AvgP {
  TView x;
  int kd, OD;
  float* out;
  long long total_warps;   // N * OD * (C / 8)
};

// One WARP per (n, od, 8-channel group): the lanes split the kd*H*W window positions (98 for the [2,7,7] window of
// i3d.py:293-294), each lane accumulates its positions' 8 channels from 16-byte loads, then five rounds of warp
// shuffles reduce the 32 partial sums.  (The previous kernel walked the 98 positions serially in one thread per
// channel group on N*OD blocks - 32 of the 148 SMs for a 32-clip batch.)
__global__ void __launch_bounds__(256) avgpool_kernel(const AvgP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c8n = p.x.C >> 3;
  const int hw = p.x.H * p.x.W, win = p.kd * hw;
  const float inv = 1.f / static_cast<float>(win);
  const long long wstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long wi = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); wi < p.total_warps; wi += wstride) {
    const int c8 = static_cast<int>(wi % c8n);
    const long long t = wi / c8n;
    const int od = static_cast<int>(t % p.OD), n = static_cast<int>(t / p.OD);
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    for (int pos = lane; pos < win; pos += 32) {
      const int a = pos / hw, r = pos - a * hw, h = r / p.x.W, w = r - h * p.x.W;
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(elem_ptr(p.x, pix_index(p.x, n, od + a, h, w), c8 * 8))), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
    if (lane == 0) {
      float4* o = reinterpret_cast<float4*>(p.out + (static_cast<long long>(n) * p.OD + od) * p.x.C + c8 * 8);
      o[0] = make_float4(s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv);
      o[1] = make_float4(s[4] * inv, s[5] * inv, s[6] * inv, s[7] * inv);
    }
  }
}

// ------------------------------------------- consumer-side view of a feature matrix (MGFN Dataset.__getitem__)
struct MgfnP {
  const float* feats;     // [T][ncrops][F] fp32
  const int32_t* bounds;  // train: [seg + 1] segment boundaries (np.linspace(0, T, seg + 1, dtype=int)); test: unused
  float* out;             // train: [ncrops][seg][F + 1]; test: [T][ncrops][F + 1]
  int T, ncrops, F, seg, train;
};

// One block per output row.  train (dataset.py:87-99 + utils.py:34-42 process_feat): row (crop c, segment s) = mean of
// snippet rows bounds[s] .. bounds[s+1]-1 of crop c (the single row bounds[s] when the segment is empty), then its L2
// magnitude as column F.  test (dataset.py:68-86): row (t, c) = the snippet row and its magnitude.  The sum of squares
// is reduced with warp shuffles, then across the block's warps through shared memory.
__global__ void __launch_bounds__(256) mgfn_rows_kernel(const MgfnP p) {
  __shared__ float warp_sums[8];
  const int row = blockIdx.x;
  int c, r0, r1;
  float* o;
  if (p.train) {
    c = row / p.seg;
    const int sgm = row - c * p.seg;
    r0 = p.bounds[sgm]; r1 = p.bounds[sgm + 1];
    if (r0 == r1) r1 = r0 + 1;
    o = p.out + static_cast<long long>(row) * (p.F + 1);
  } else {
    const int t = row / p.ncrops;
    c = row - t * p.ncrops;
    r0 = t; r1 = t + 1;
    o = p.out + static_cast<long long>(row) * (p.F + 1);
  }
  const float inv = 1.f / static_cast<float>(r1 - r0);
  float sq = 0.f;
  for (int f = threadIdx.x; f < p.F; f += blockDim.x) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc += __ldg(p.feats + (static_cast<long long>(r) * p.ncrops + c) * p.F + f);
    const float v = (r1 - r0) > 1 ? acc * inv : acc;
    o[f] = v;
    sq = fmaf(v, v, sq);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += warp_sums[w];
    o[p.F] = sqrtf(tot);
  }
}

// ----------------------------------------------------------------- preprocess
// Crop + resize + normalise uint8 frames into the anonymizer's bf16 input (dali_extraction.py:38-50 /
// shanghai_dl.py:27-40).  One block = one band of BH output rows of one output image; a thread owns one output COLUMN:
//   1. every thread computes the resampling entry of its column in registers (start, KX weights), the first BH threads
//      those of the band's rows (shared memory);
//   2. the source rows the band needs are staged as raw bytes with coalesced 16-byte loads;
//   3. horizontal pass: for every staged row the thread interpolates its column (KX taps x 3 channels; AA_FLOAT:
//      fp32 FMAs over the bytes in torchvision's tap order, the /255 applied to the sum; PIL_U8: Pillow's 22-bit
//      fixed point, rounded to 8 bits) into a [row][channel][column] plane in shared memory;
//   4. vertical pass: the thread walks down its column, KY taps per output row, and writes ONE 16- or 8-byte pixel
//      (3 real channels + zero padding to the 8 channels the UNet stem / 4 channels the 7x7 stems read).
// The separable form does each horizontal interpolation once per source row instead of once per output row that
// touches it, there is no integer division and no table lookup in the loops, and tap counts are compile-time
// (weights beyond a column's real tap count are zero, their indices clamped): ncu on the round-1 kernel and on the
// first staged version showed both ISSUE-bound (86 % issue-active, 345 M warp instructions for 25.7 M pixels).
constexpr int PP_KMAX = 8;
constexpr int PP_MAXOUT = 4096;   // (the band height shrinks until the staged rows + planes fit in shared memory)
constexpr int PP_THREADS = 224;   // = the 224 output columns of the hot configuration: no idle thread in the compute phase

struct PrepP {
  const uint8_t* frames;
  long long frames_bytes;
  int F, Hs, Ws;
  const int32_t* desc;
  int n_out, crop_h, crop_w, resample;
  TView y;
  float* frames_f32;
  int BH, max_rows, pitch;   // output rows per block; staged source rows (bound); staged row pitch in bytes
  float sx, supx, invx;      // AA_FLOAT x axis: scale = crop_w / Wo, support = max(scale, 1), 1 / support (host floats)
};
constexpr int PP_FAST_BH = 8;   // the FAST instantiations: bands of exactly 8 rows, vector stores, no fp32 copy

// One entry of a resampling table: first tap, tap count, K weights (zero beyond the count).
// AA_FLOAT follows aten's _upsample_bilinear2d_aa (align_corners=False): support = max(scale,1),
// taps j in [lo,hi) with w = 1 - |(j - center + 0.5)/support|, normalised.
// PIL_U8 follows Pillow's precompute_coeffs + normalize_coeffs_8bpc (double weights -> 22-bit fixed).
template <int K>
__device__ __forceinline__ void axis_entry(int in, int out, int resample, int i, int& lo, int& cnt, float (&wf)[K], int (&wi)[K]) {
#pragma unroll
  for (int x = 0; x < K; ++x) { wf[x] = 0.f; wi[x] = 0; }
  if (resample == TEDSPAD_RESAMPLE_PIL_U8) {
    const double scale = static_cast<double>(in) / out;
    const double fscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * fscale;
    const double center = (i + 0.5) * scale;
    const double ss = 1.0 / fscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in) xmax = in;
    xmax -= xmin;
    if (xmax > K) xmax = K;
    double k[K];
    double ww = 0.0;
#pragma unroll
    for (int x = 0; x < K; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = (x < xmax && a < 1.0) ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
#pragma unroll
    for (int x = 0; x < K; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * static_cast<double>(1 << 22);
      wi[x] = x < xmax ? static_cast<int>(k[x] < 0 ? -0.5 + v : 0.5 + v) : 0;
    }
    lo = xmin;
    cnt = xmax;
  } else {
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    const float support = scale >= 1.f ? scale : 1.f;
    const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float center = scale * (i + 0.5f);
    int xmin = static_cast<int>(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5f);
    if (xmax > in) xmax = in;
    int n = xmax - xmin;
    if (n > K) n = K;
    float tot = 0.f;
#pragma unroll
    for (int x = 0; x < K; ++x) {
      float a = (x + xmin - center + 0.5f) * invscale;
      if (a < 0.f) a = -a;
      const float w = (x < n && a < 1.f) ? 1.f - a : 0.f;
      wf[x] = w;
      if (x < n) tot += w;       // (summed in tap order, like aten)
    }
#pragma unroll
    for (int x = 0; x < K; ++x)
      if (tot != 0.f && x < n && wf[x] != 0.f) wf[x] /= tot;   // (0 / tot = 0, but through the division's slow path)
    lo = xmin;
    cnt = n;
  }
}

__device__ __forceinline__ int clip8_fixed(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// b / 255.f, correctly rounded, without a division: q = b*r, one Newton step on the exact remainder (r = rn(1/255);
// checked for all 256 bytes against IEEE division with exact rational arithmetic, tests/test_host.py)
__device__ __forceinline__ float u8_over_255(uint32_t b) {
  const float x = static_cast<float>(b), r = __uint_as_float(0x3b808081u);
  const float q = x * r;
  return fmaf(fmaf(-q, 255.f, x), r, q);
}

// float(b) for a byte: 2^23 + b is exact in fp32 (OR the byte into the mantissa of 2^23), minus 2^23
__device__ __forceinline__ float u8_to_float(uint32_t b) { return __uint_as_float(0x4b000000u | b) - 8388608.f; }

// FAST = 0: any band height / output layout / optional fp32 copy.  FAST = 4 | 8: the hot configuration - bands of
// PP_FAST_BH full rows, pixels of FAST channels stored with one vector store, no fp32 copy: the vertical loop unrolls
// completely and its addressing constant-folds (the generic loop spent 50 instructions per pixel, half of them uniform
// bookkeeping).
template <bool PIL, int KX, int KY, int FAST>
__global__ void __launch_bounds__(PP_THREADS, (FAST != 0 && !PIL && KY == 2) ? 7 : 1) preprocess_kernel(const PrepP p) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int Ho = p.y.H, Wo = p.y.W, BH = FAST ? PP_FAST_BH : p.BH;
  int* ylo = reinterpret_cast<int*>(pp_smem);                   // [BH + 1]: first source row per output row, then the band's end
  int* rsh = ylo + BH + 1;                                      // byte shift of every staged row
  float* ywf = reinterpret_cast<float*>(rsh + p.max_rows);      // [KY][BH]
  int* ywi = reinterpret_cast<int*>(ywf);
  // (offsets computed in words: a pointer -> integer cast would lose the shared address space and make the compiler
  // rebuild the shared-window base inside every loop)
  const int raw_off = ((BH + 1 + p.max_rows + KY * BH + 3) & ~3) * 4;
  uint8_t* raw = pp_smem + raw_off;                                                          // 16-byte aligned
  float* hf = reinterpret_cast<float*>(pp_smem + raw_off + p.max_rows * p.pitch);            // [row][channel][column]
  uint8_t* hb = pp_smem + raw_off + p.max_rows * p.pitch;

  const int n = blockIdx.y;
  const int oy0 = blockIdx.x * BH, nrow = FAST ? PP_FAST_BH : min(BH, Ho - oy0);
  if (threadIdx.x < nrow) {
    int lo, cnt, wi[KY];
    float wf[KY];
    axis_entry<KY>(p.crop_h, Ho, p.resample, oy0 + threadIdx.x, lo, cnt, wf, wi);
    ylo[threadIdx.x] = lo;
#pragma unroll
    for (int a = 0; a < KY; ++a) {
      if (PIL) ywi[a * BH + threadIdx.x] = wi[a]; else ywf[a * BH + threadIdx.x] = wf[a];
    }
    if (threadIdx.x == nrow - 1) ylo[BH] = lo + cnt;            // end of the band's source rows (slot BH of ylo)
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();   // the tables above depend on the launch parameters only

  const int32_t* d = p.desc + n * 4;
  const int src = d[0], top = d[1], left = d[2], flip = d[3];
  const int r0 = ylo[0];
  const int R = min(ylo[BH] - r0, p.max_rows);
  const int nb = p.crop_w * 3;                                  // bytes of one cropped source row
  if (src >= 0) {
    // stage rows [top + r0, top + r0 + R) x columns [col0, col0 + crop_w) (the mirrored window when flipped), raw bytes
    const int col0 = flip ? p.Ws - left - p.crop_w : left;
    const uint8_t* fend = p.frames + p.frames_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < R; r += PP_THREADS / 32) {          // a warp per source row: the row address once per warp
      const uint8_t* g = p.frames + ((static_cast<long long>(src) * p.Hs + top + r0 + r) * p.Ws + col0) * 3;
      const int shift = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
      const uint8_t* a0 = g - shift;
      const int nchunk = (shift + nb + 15) >> 4;
      uint8_t* dst = raw + r * p.pitch;
      if (lane == 0) rsh[r] = shift;
      for (int i = lane; i < nchunk; i += 32) {
        const uint8_t* a = a0 + 16 * i;
        uint4 q;
        if (a + 16 <= fend) {
          q = __ldg(reinterpret_cast<const uint4*>(a));
        } else {   // last bytes of the frame buffer: no read past its end
          uint8_t* qb = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
          for (int k = 0; k < 16; ++k) qb[k] = (a + k < fend) ? a[k] : 0;
        }
        *reinterpret_cast<uint4*>(dst + 16 * i) = q;
      }
    }
  }
  __syncthreads();

  const float inv255 = 1.f / (255.f * 4194304.f);   // AA: pixels are /255 (dali_extraction.py:41) and the horizontal
                                                    // weights are 22-bit fixed point: both applied to the sums
  constexpr bool STREAM = FAST != 0 && !PIL && KY == 2;
  const int cpy = p.y.C;   // 8 (UNet stem) or 4 (7x7 stems) when vectorisable
  const bool vec8 = cpy == 8 && ((p.y.ld | p.y.coff) & 7) == 0, vec4 = cpy == 4 && ((p.y.ld | p.y.coff) & 3) == 0;
  for (int ox = threadIdx.x; ox < Wo; ox += PP_THREADS) {
    if (src >= 0) {
      // ---- horizontal pass: this column of every staged row
      int xl, xn, xwi[KX];
      if (PIL) {
        float xwf[KX];
        axis_entry<KX>(p.crop_w, Wo, p.resample, ox, xl, xn, xwf, xwi);
      } else {
        // aten's antialias weights of this column (same float arithmetic as axis_entry; scale / support / 1 / support
        // come from the host), normalised and quantised to 22-bit fixed point with ONE division
        const float center = p.sx * (ox + 0.5f);
        xl = max(static_cast<int>(center - p.supx + 0.5f), 0);
        xn = min(min(static_cast<int>(center + p.supx + 0.5f), p.crop_w) - xl, KX);
        float w[KX], tot = 0.f;
#pragma unroll
        for (int b = 0; b < KX; ++b) {
          const float a = fabsf((b + xl - center + 0.5f) * p.invx);
          w[b] = (b < xn && a < 1.f) ? 1.f - a : 0.f;
          tot += w[b];
        }
        const float q = tot != 0.f ? 4194304.f / tot : 0.f;
#pragma unroll
        for (int b = 0; b < KX; ++b) xwi[b] = static_cast<int>(w[b] * q + 0.5f);
      }
      int off[KX];   // byte offset of tap b inside a staged row (mirrored when the crop comes from the flipped frame)
#pragma unroll
      for (int b = 0; b < KX; ++b) {
        const int col = min(xl + b, p.crop_w - 1);
        off[b] = (flip ? p.crop_w - 1 - col : col) * 3;
      }
      if (STREAM) {
        // STREAMING form of the two passes (2 vertical taps: every up-scaling axis): the thread walks down its column
        // with the horizontally interpolated source rows yl and yl + 1 of the current output row in registers; the next
        // output row starts at the same source row (both re-used), the next one (one new row) or further down (two),
        // which is the same for every thread of the block (uniform branches).  No plane in shared memory: 3 stores per
        // source row and 6 loads + their addressing per output pixel less; the values and the order of the two FMAs
        // are those of the generic path, bit for bit.
        auto hrow = [&](int r, float& v0, float& v1, float& v2) {
          const uint8_t* rowp = raw + r * p.pitch + rsh[r];
          int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
          for (int b = 0; b < KX; ++b) {
            const uint8_t* px = rowp + off[b];
            h0 += px[0] * xwi[b]; h1 += px[1] * xwi[b]; h2 += px[2] * xwi[b];
          }
          v0 = static_cast<float>(h0) * inv255; v1 = static_cast<float>(h1) * inv255; v2 = static_cast<float>(h2) * inv255;
        };
        __nv_bfloat16* yq = elem_ptr_w(p.y, pix_index(p.y, n, 0, oy0, ox), 0);
        const int y_row = p.y.Wp * p.y.ld;
        const int rmax = R - 1;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
        int base = -4;
#pragma unroll
        for (int ly = 0; ly < PP_FAST_BH; ++ly, yq += y_row) {
          const int yl = ylo[ly] - r0;
          if (yl == base + 1) {
            a0 = b0; a1 = b1; a2 = b2;
            hrow(min(yl + 1, rmax), b0, b1, b2);
          } else if (yl != base) {
            hrow(yl, a0, a1, a2);
            hrow(min(yl + 1, rmax), b0, b1, b2);
          }
          base = yl;
          const float w0 = ywf[ly], w1 = ywf[BH + ly];
          const float o0 = fmaf(w1, b0, fmaf(w0, a0, 0.f)), o1 = fmaf(w1, b1, fmaf(w0, a1, 0.f)), o2 = fmaf(w1, b2, fmaf(w0, a2, 0.f));
          const uint32_t lo = cvt_bf16x2(o0, o1, false), hi = cvt_bf16x2(o2, 0.f, false);
          if (FAST == 8) *reinterpret_cast<uint4*>(yq) = make_uint4(lo, hi, 0u, 0u);
          else *reinterpret_cast<uint2*>(yq) = make_uint2(lo, hi);
        }
      } else
      for (int r = 0; r < R; ++r) {
        const uint8_t* rowp = raw + r * p.pitch + rsh[r];
        if (PIL) {
          int h0 = 1 << 21, h1 = 1 << 21, h2 = 1 << 21;
#pragma unroll
          for (int b = 0; b < KX; ++b) {
            const uint8_t* px = rowp + off[b];
            h0 += px[0] * xwi[b]; h1 += px[1] * xwi[b]; h2 += px[2] * xwi[b];
          }
          uint8_t* hp = hb + (r * 3) * Wo + ox;
          hp[0] = static_cast<uint8_t>(clip8_fixed(h0)); hp[Wo] = static_cast<uint8_t>(clip8_fixed(h1));
          hp[2 * Wo] = static_cast<uint8_t>(clip8_fixed(h2));
        } else {
          // sum of w_b * byte_b with the weights in 22-bit fixed point (|dw| <= 2^-23: 4e-7 on a [0,1] pixel, two
          // orders below the 3e-5 the float kernel of aten itself is from exact arithmetic): one IMAD per byte, no
          // byte -> float conversion; the sum (< 2^30) becomes a float once per channel, scaled by 1 / (255 * 2^22)
          int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
          for (int b = 0; b < KX; ++b) {
            const uint8_t* px = rowp + off[b];
            h0 += px[0] * xwi[b]; h1 += px[1] * xwi[b]; h2 += px[2] * xwi[b];
          }
          float* hp = hf + (r * 3) * Wo + ox;
          hp[0] = static_cast<float>(h0) * inv255; hp[Wo] = static_cast<float>(h1) * inv255;
          hp[2 * Wo] = static_cast<float>(h2) * inv255;
        }
      }
    }
    if (STREAM && src >= 0) continue;   // (written by the streaming form above)
    // ---- vertical pass: down the column (only this thread wrote / reads column ox of the planes: no barrier needed)
    __nv_bfloat16* yp = elem_ptr_w(p.y, pix_index(p.y, n, 0, oy0, ox), 0);
    const int y_row = p.y.Wp * p.y.ld;                          // elements per output row (< 2^31: PP_MAXOUT x ld)
    const long long plane = static_cast<long long>(Ho) * Wo;
    float* fo = p.frames_f32 ? p.frames_f32 + static_cast<long long>(n) * 3 * plane + static_cast<long long>(oy0) * Wo + ox : nullptr;
    const bool f32 = !FAST && fo != nullptr;
    const int hstride = 3 * Wo, rmax = R - 1;
    const float* hcol = hf + ox;
    const uint8_t* bcol = hb + ox;
#pragma unroll
    for (int ly = 0; ly < (FAST ? PP_FAST_BH : nrow); ++ly) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      if (src >= 0) {
        const int yl = ylo[ly] - r0;
        if (PIL) {
          int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
          for (int a = 0; a < KY; ++a) {
            const uint8_t* hp = bcol + min(yl + a, rmax) * hstride;
            const int ky = ywi[a * BH + ly];
            a0 += hp[0] * ky; a1 += hp[Wo] * ky; a2 += hp[2 * Wo] * ky;
          }
          o0 = u8_over_255(static_cast<uint32_t>(clip8_fixed(a0)));
          o1 = u8_over_255(static_cast<uint32_t>(clip8_fixed(a1)));
          o2 = u8_over_255(static_cast<uint32_t>(clip8_fixed(a2)));
        } else {
#pragma unroll
          for (int a = 0; a < KY; ++a) {
            const float* hp = hcol + min(yl + a, rmax) * hstride;
            const float wy = ywf[a * BH + ly];
            o0 = fmaf(wy, hp[0], o0); o1 = fmaf(wy, hp[Wo], o1); o2 = fmaf(wy, hp[2 * Wo], o2);
          }
        }
      }
      const uint32_t lo = cvt_bf16x2(o0, o1, false), hi = cvt_bf16x2(o2, 0.f, false);
      __nv_bfloat16* yq = yp + ly * y_row;
      if (FAST == 8 || (!FAST && vec8)) {
        *reinterpret_cast<uint4*>(yq) = make_uint4(lo, hi, 0u, 0u);
      } else if (FAST == 4 || (!FAST && vec4)) {
        *reinterpret_cast<uint2*>(yq) = make_uint2(lo, hi);
      } else {
        const float o[3] = {o0, o1, o2};
        for (int c = 0; c < cpy; ++c) yq[c] = __float2bfloat16_rn(c < 3 ? o[c] : 0.f);
      }
      if (f32) {
        float* fq = fo + ly * Wo;
        fq[0] = o0; fq[plane] = o1; fq[2 * plane] = o2;
      }
    }
  }
}

// --------------------------------------------------- fp32 NC(D)HW -> bf16 channels-last
struct CvtP {
  const float* x;
  int Cx;
  TView y;
  long long total;  // N*D*H*W*(y.C/8)
};

__global__ void __launch_bounds__(256) nchw_to_cl_kernel(const CvtP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const long long plane = static_cast<long long>(p.y.D) * p.y.H * p.y.W;
  const int cw = p.y.C == 4 ? 4 : 8;  // channels per thread: one 8- or 16-byte store
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < p.total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    // pixel fastest so that the fp32 plane reads are coalesced
    const long long pixl = idx % (plane * p.y.N);
    const int cg = static_cast<int>(idx / (plane * p.y.N));
    const int n = static_cast<int>(pixl / plane);
    long long r = pixl - n * plane;
    const int w = static_cast<int>(r % p.y.W); r /= p.y.W;
    const int h = static_cast<int>(r % p.y.H);
    const int d = static_cast<int>(r / p.y.H);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cg * cw + i;
      f[i] = (i < cw && c < p.Cx) ? __ldg(p.x + (static_cast<long long>(n) * p.Cx + c) * plane + (pixl - n * plane)) : 0.f;
    }
    __nv_bfloat16* yp = elem_ptr_w(p.y, pix_index(p.y, n, d, h, w), cg * cw);
    const uint4 q = pack8(f);
    if (cw == 4) *reinterpret_cast<uint2*>(yp) = make_uint2(q.x, q.y);
    else *reinterpret_cast<uint4*>(yp) = q;
  }
}

// ------------------------------------- planar anonymizer output -> encoder clip (raw-reshape glue)
struct P2CP {
  const __nv_bfloat16* planes;  // [B*T][3][H][W]
  TView y;                      // [B][T][H][W][C>=3], C in {4, 8}
  int T;
  long long total;              // B*T*H*(W/8)
};

// thread -> 8 consecutive pixels of one encoder row: three 16-byte plane reads, 8 pixel stores
__global__ void __launch_bounds__(256) planes_to_clip_kernel(const P2CP p) {
  pdl_launch_dependents();   // programmatic dependent launch, see common.h
  pdl_wait();
  const int w8n = p.y.W >> 3;
  const long long plane = static_cast<long long>(p.y.H) * p.y.W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < p.total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = idx;
    const int w8 = static_cast<int>(t % w8n); t /= w8n;
    const int h = static_cast<int>(t % p.y.H); t /= p.y.H;
    const int te = static_cast<int>(t % p.T);
    const int b = static_cast<int>(t / p.T);
    uint4 src[3];
#pragma unroll
    for (int ce = 0; ce < 3; ++ce) {
      // encoder channel ce, time te  <-  plane ce*T + te of the clip's 3T planes (dali_extraction.py:173)
      const long long pl = static_cast<long long>(b) * 3 * p.T + ce * p.T + te;
      src[ce] = __ldg(reinterpret_cast<const uint4*>(p.planes + pl * plane + static_cast<long long>(h) * p.y.W + w8 * 8));
    }
    const uint16_t* s0 = reinterpret_cast<const uint16_t*>(&src[0]);
    const uint16_t* s1 = reinterpret_cast<const uint16_t*>(&src[1]);
    const uint16_t* s2 = reinterpret_cast<const uint16_t*>(&src[2]);
    __nv_bfloat16* yp = elem_ptr_w(p.y, pix_index(p.y, b, te, h, w8 * 8), 0);
    if (p.y.C == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint2 o;
        o.x = static_cast<uint32_t>(s0[i]) | (static_cast<uint32_t>(s1[i]) << 16);
        o.y = static_cast<uint32_t>(s2[i]);
        *reinterpret_cast<uint2*>(yp + static_cast<long long>(i) * p.y.ld) = o;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 o;
        o.x = static_cast<uint32_t>(s0[i]) | (static_cast<uint32_t>(s1[i]) << 16);
        o.y = static_cast<uint32_t>(s2[i]);
        o.z = 0u; o.w = 0u;
        *reinterpret_cast<uint4*>(yp + static_cast<long long>(i) * p.y.ld) = o;
      }
    }
  }
}

// ------------------------------------------------------------------ nearest x2 into a channel slice
struct NearP {
  TView x, y;
  uint32_t c8_magic;
  long long total;  // input rows N*H
};

// One block per INPUT row: a thread loads one 16-byte channel chunk of an input pixel once and stores it to the 2x2
// output pixels it covers (F.interpolate(scale_factor=2, mode="nearest"): out[y][x] = in[y/2][x/2]).
__global__ void __launch_bounds__(256) upsample2x_nearest_kernel(const NearP p) {
  pdl_launch_dependents();
  pdl_wait();
  const int c8n = p.x.C >> 3;
  const int items = p.x.W * c8n;
  const long long y_row = static_cast<long long>(p.y.Wp) * p.y.ld;
  for (long long row = blockIdx.x; row < p.total; row += gridDim.x) {
    const int n = static_cast<int>(row / p.x.H), ih = static_cast<int>(row - static_cast<long long>(n) * p.x.H);
    const __nv_bfloat16* xr = elem_ptr(p.x, pix_index(p.x, n, 0, ih, 0), 0);
    __nv_bfloat16* yr = elem_ptr_w(p.y, pix_index(p.y, n, 0, 2 * ih, 0), 0);
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int iw = fast_div(it, p.c8_magic), c8 = it - iw * c8n;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(xr + static_cast<long long>(iw) * p.x.ld + c8 * 8));
      __nv_bfloat16* yp = yr + static_cast<long long>(2 * iw) * p.y.ld + c8 * 8;
      *reinterpret_cast<uint4*>(yp) = q;
      *reinterpret_cast<uint4*>(yp + p.y.ld) = q;
      *reinterpret_cast<uint4*>(yp + y_row) = q;
      *reinterpret_cast<uint4*>(yp + y_row + p.y.ld) = q;
    }
  }
}

// ------------------------------------- channels-last anonymizer output -> encoder clip (raw-reshape glue)
struct F2CP {
  TView x;          // [B*T][1][H][W][>=3] bf16 (the UNet++ head output: unbounded, activation=None)
  TView y;          // [B][T][H][W][4|8]
  float* frames;    // optional fp32 [B*T][3][H][W]
  int T;
  int s2d;          // x is [B*T][1][H/2][W/2][>=12]: channel (2*(h&1) + (w&1))*3 + c of pixel (h/2, w/2)
  long long total;  // B*T*H*W clip pixels
};

// thread -> one ENCODER pixel (b, te, h, w): channel ce comes from plane ce*T + te of the clip's 3T anonymizer planes,
// i.e. colour (ce*T + te) % 3 of frame (ce*T + te) / 3 (dali_extraction.py:171-173); one 8- or 16-byte store.
__global__ void __launch_bounds__(256) frames_to_clip_kernel(const F2CP p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long plane = static_cast<long long>(p.y.H) * p.y.W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < p.total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = idx;
    const int w = static_cast<int>(t % p.y.W); t /= p.y.W;
    const int h = static_cast<int>(t % p.y.H); t /= p.y.H;
    const int te = static_cast<int>(t % p.T);
    const int b = static_cast<int>(t / p.T);
    uint16_t v[3];
    const int xh = p.s2d ? h >> 1 : h, xw = p.s2d ? w >> 1 : w, xc = p.s2d ? (2 * (h & 1) + (w & 1)) * 3 : 0;
#pragma unroll
    for (int ce = 0; ce < 3; ++ce) {
      const int pl = ce * p.T + te, tf = pl / 3, o = pl - 3 * tf;
      v[ce] = __ldg(reinterpret_cast<const uint16_t*>(elem_ptr(p.x, pix_index(p.x, b * p.T + tf, 0, xh, xw), xc + o)));
    }
    __nv_bfloat16* yp = elem_ptr_w(p.y, pix_index(p.y, b, te, h, w), 0);
    const uint32_t lo = static_cast<uint32_t>(v[0]) | (static_cast<uint32_t>(v[1]) << 16), hi = static_cast<uint32_t>(v[2]);
    if (p.y.C == 4) *reinterpret_cast<uint2*>(yp) = make_uint2(lo, hi);
    else *reinterpret_cast<uint4*>(yp) = make_uint4(lo, hi, 0u, 0u);
    if (p.frames != nullptr) {
      // the un-scattered frames (the fa_model return value): frame b*T + te, its own three colours
      const __nv_bfloat16* xp = elem_ptr(p.x, pix_index(p.x, b * p.T + te, 0, xh, xw), xc);
      float* fo = p.frames + (static_cast<long long>(b * p.T + te) * 3) * plane + static_cast<long long>(h) * p.y.W + w;
      fo[0] = __bfloat162float(xp[0]); fo[plane] = __bfloat162float(xp[1]); fo[2 * plane] = __bfloat162float(xp[2]);
    }
  }
}

// ------------------------------------------------------------------ row-wise L2 normalisation (fp32, in place)
struct NormP {
  float* x;
  int rows, cols;
  float eps;
};

// nn.functional.normalize(x, p=2, dim=1) (aux_code/model_loaders.py:252): x / max(||x||_2, eps); one block per row,
// warp-shuffle + shared-memory reduction of the sum of squares
__global__ void __launch_bounds__(128) l2_normalize_rows_kernel(const NormP p) {
  __shared__ float warp_sums[4];
  pdl_launch_dependents();
  pdl_wait();
  float* row = p.x + static_cast<long long>(blockIdx.x) * p.cols;
  float sq = 0.f;
  for (int i = threadIdx.x; i < p.cols; i += blockDim.x) sq = fmaf(row[i], row[i], sq);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sq;
  __syncthreads();
  const float inv = 1.f / fmaxf(sqrtf(warp_sums[0] + warp_sums[1] + warp_sums[2] + warp_sums[3]), p.eps);
  for (int i = threadIdx.x; i < p.cols; i += blockDim.x) row[i] *= inv;
}

static int grid_for(long long total, int threads) {
  const long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// grid for the row-per-block kernels: every row when rows are long enough to fill a block, else fewer
// blocks that loop over rows (so that tiny rows do not launch mostly idle blocks)
static int rows_grid(long long rows, int items_per_row) {
  const long long cap = static_cast<long long>(num_sms()) * 64;
  long long blocks = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  if (items_per_row < 128) blocks = (rows * items_per_row + 255) / 256;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks > cap ? cap : (blocks > rows ? rows : blocks));
}

}  // namespace tsp

using namespace tsp;

extern "C" int tedspad_maxpool(const tedspad_tensor* x, const tedspad_tensor* y, int32_t kd, int32_t kh, int32_t kw,
                               int32_t sd, int32_t sh, int32_t sw, int32_t pd, int32_t ph, int32_t pw, int32_t zero_pad,
                               void* stream) {
  TSP_CHECK(x && y, "maxpool: null tensor");
  if (check_tensor(*x, "maxpool.x", 8) || check_tensor(*y, "maxpool.y", 8)) return 1;
  TSP_CHECK(x->C == y->C && x->C % 8 == 0 && x->N == y->N, "maxpool: channel/batch mismatch");
  TSP_CHECK(kd >= 1 && kh >= 1 && kw >= 1 && sd >= 1 && sh >= 1 && sw >= 1, "maxpool: bad window");
  TSP_CHECK((y->D - 1) * sd - pd < x->D && (y->H - 1) * sh - ph < x->H && (y->W - 1) * sw - pw < x->W,
            "maxpool: output extents exceed input");
  PoolP p;
  p.x = make_view(*x); p.y = make_view(*y);
  p.kd = kd; p.kh = kh; p.kw = kw; p.sd = sd; p.sh = sh; p.sw = sw; p.pd = pd; p.ph = ph; p.pw = pw;
  p.zero_pad = zero_pad;
  p.c8_magic = static_cast<uint32_t>((0x100000000ULL + (y->C / 8) - 1) / (y->C / 8));
  TSP_CHECK(static_cast<long long>(y->W) * (y->C / 8) < 65536, "maxpool: row of %d x %d channels too long", y->W, y->C);
  p.total = static_cast<long long>(y->N) * y->D * y->H;   // output rows
  TSP_CHECK(p.total < (1LL << 31), "maxpool: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = rows_grid(p.total, y->W * (y->C / 8));
  const bool same333 = kd == 3 && kh == 3 && kw == 3 && sd == 1 && sh == 1 && sw == 1 && pd == 1 && ph == 1 && pw == 1 &&
                       x->D == y->D && x->H == y->H && x->W == y->W;
  if (same333) {
    static const bool rb = [] { const char* e = getenv("TEDSPAD_POOL_RB"); return e == nullptr || e[0] != '0'; }();
    if (rb) {
      const long long items = static_cast<long long>(y->N) * ((y->D + 1) / 2) * ((y->H + 1) / 2) * (y->C / 8);
      TSP_CUDA(launch_kernel(maxpool333_s1_rb_kernel, dim3(grid_for(items, 256)), dim3(256), 0, st, p));
    } else {
      TSP_CUDA(launch_kernel(maxpool333_s1_kernel, dim3(grid_for(p.total * (y->C / 8), 256)), dim3(256), 0, st, p));
    }
  } else if (kd == 1 && kh == 3 && kw == 3) {
    TSP_CUDA(launch_kernel(maxpool_kernel<1, 3, 3>, dim3(grid), dim3(256), 0, st, p));
  } else if (kd == 3 && kh == 3 && kw == 3) {
    TSP_CUDA(launch_kernel(maxpool_kernel<3, 3, 3>, dim3(grid), dim3(256), 0, st, p));
  } else if (kd == 2 && kh == 2 && kw == 2) {
    TSP_CUDA(launch_kernel(maxpool_kernel<2, 2, 2>, dim3(grid), dim3(256), 0, st, p));
  } else if (kd == 1 && kh == 2 && kw == 2) {
    TSP_CUDA(launch_kernel(maxpool_kernel<1, 2, 2>, dim3(grid), dim3(256), 0, st, p));
  } else if (kd == 2 && kh == 3 && kw == 3) {
    TSP_CUDA(launch_kernel(maxpool_kernel<2, 3, 3>, dim3(grid), dim3(256), 0, st, p));
  } else if (kd == 2 && kh == 1 && kw == 1) {
    TSP_CUDA(launch_kernel(maxpool_kernel<2, 1, 1>, dim3(grid), dim3(256), 0, st, p));
  } else {
    TSP_CUDA(launch_kernel(maxpool_kernel<0, 0, 0>, dim3(grid), dim3(256), 0, st, p));
  }
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_upsample2x(const tedspad_tensor* x, const tedspad_tensor* y, void* stream) {
  TSP_CHECK(x && y, "upsample2x: null tensor");
  if (check_tensor(*x, "upsample2x.x", 8) || check_tensor(*y, "upsample2x.y", 8)) return 1;
  TSP_CHECK(x->C == y->C && x->C % 8 == 0 && x->N == y->N && x->D == 1 && y->D == 1, "upsample2x: shape mismatch");
  UpP p;
  p.x = make_view(*x); p.y = make_view(*y);
  p.UH = 2 * x->H; p.UW = 2 * x->W;
  TSP_CHECK(y->H >= p.UH && y->W >= p.UW, "upsample2x: target smaller than 2x source");
  p.offy = (y->H - p.UH) / 2; p.offx = (y->W - p.UW) / 2;  // F.pad(diff//2, diff - diff//2)
  p.sy = p.UH > 1 ? static_cast<float>(x->H - 1) / static_cast<float>(p.UH - 1) : 0.f;
  p.sx = p.UW > 1 ? static_cast<float>(x->W - 1) / static_cast<float>(p.UW - 1) : 0.f;
  p.total = static_cast<long long>(y->N) * ((y->H + UP_ROWS - 1) / UP_ROWS);   // bands of UP_ROWS output rows
  p.c8_magic = static_cast<uint32_t>((0x100000000ULL + (y->C / 8) - 1) / (y->C / 8));
  TSP_CHECK(static_cast<long long>(y->W) * (y->C / 8) < 65536, "upsample2x: row of %d x %d channels too long", y->W, y->C);
  TSP_CUDA(launch_kernel(upsample2x_kernel, dim3(rows_grid(p.total, y->W * (y->C / 8))), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_upsample2x_nearest(const tedspad_tensor* x, const tedspad_tensor* y, void* stream) {
  TSP_CHECK(x && y, "upsample2x_nearest: null tensor");
  if (check_tensor(*x, "upsample2x_nearest.x", 8) || check_tensor(*y, "upsample2x_nearest.y", 8)) return 1;
  TSP_CHECK(x->C == y->C && x->C % 8 == 0 && x->N == y->N && x->D == 1 && y->D == 1 && y->H == 2 * x->H && y->W == 2 * x->W,
            "upsample2x_nearest: [%d,%d,%d,%d] -> [%d,%d,%d,%d] is not an exact x2", x->N, x->H, x->W, x->C, y->N, y->H,
            y->W, y->C);
  NearP p;
  p.x = make_view(*x); p.y = make_view(*y);
  p.c8_magic = static_cast<uint32_t>((0x100000000ULL + (x->C / 8) - 1) / (x->C / 8));
  TSP_CHECK(static_cast<long long>(x->W) * (x->C / 8) < 65536, "upsample2x_nearest: row of %d x %d channels too long", x->W, x->C);
  p.total = static_cast<long long>(x->N) * x->H;
  TSP_CUDA(launch_kernel(upsample2x_nearest_kernel, dim3(rows_grid(p.total, x->W * (x->C / 8))), dim3(256), 0,
                         reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_frames_to_clip(const tedspad_tensor* x, const tedspad_tensor* y, int32_t T, int32_t s2d,
                                      float* frames_out, void* stream) {
  TSP_CHECK(x && y, "frames_to_clip: null argument");
  if (check_tensor(*x, "frames_to_clip.x", 1) || check_tensor(*y, "frames_to_clip.y", 4)) return 1;
  const int f = s2d ? 2 : 1;
  TSP_CHECK(x->C >= (s2d ? 12 : 3) && x->D == 1 && (y->C == 4 || y->C == 8) && y->ld % y->C == 0 && y->coff % y->C == 0,
            "frames_to_clip: x needs >= %d channels, y must be a [B,T,H,W,4|8] clip", s2d ? 12 : 3);
  TSP_CHECK(T >= 1 && x->N % T == 0 && y->N == x->N / T && y->D == T && y->H == f * x->H && y->W == f * x->W,
            "frames_to_clip: clip [%d,%d,%d,%d] does not match %d frames of T=%d", y->N, y->D, y->H, y->W, x->N, T);
  F2CP p;
  p.x = make_view(*x); p.y = make_view(*y);
  p.frames = frames_out; p.T = T; p.s2d = s2d ? 1 : 0;
  p.total = static_cast<long long>(x->N) * y->H * y->W;
  TSP_CUDA(launch_kernel(frames_to_clip_kernel, dim3(grid_for(p.total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_outconv_sigmoid(const tedspad_tensor* x, const float* w, const float* b,
                                       const tedspad_tensor* y, int32_t T, float* frames_out, void* stream) {
  TSP_CHECK(x && y && w && b, "outconv: null argument");
  if (check_tensor(*x, "outconv.x", 8) || check_tensor(*y, "outconv.y", 1)) return 1;
  TSP_CHECK(x->C % 8 == 0 && x->D == 1, "outconv: x must be 2-D with C %% 8 == 0");
  TSP_CHECK(T >= 1 && x->N % T == 0 && y->N == x->N / T && y->D == T && y->H == x->H && y->W == x->W && y->C >= 3,
            "outconv: encoder input view [%d,%d,%d,%d,%d] does not match %d frames of T=%d", y->N, y->D, y->H, y->W,
            y->C, x->N, T);
  OutP p;
  p.x = make_view(*x); p.y = make_view(*y);
  p.w = w; p.b = b; p.frames = frames_out; p.T = T;
  p.total = static_cast<long long>(x->N) * x->H * x->W;
  const int blocks = grid_for(p.total * 8, 256);
  TSP_CUDA(launch_kernel(outconv_sigmoid_kernel, dim3(blocks), dim3(256), 3 * x->C * sizeof(float), reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_avgpool_features(const tedspad_tensor* x, int32_t kd, float* out, void* stream) {
  TSP_CHECK(x && out, "avgpool: null argument");
  if (check_tensor(*x, "avgpool.x", 8)) return 1;
  TSP_CHECK(x->C % 8 == 0, "avgpool: C %% 8 != 0");
  AvgP p;
  p.x = make_view(*x);
  p.kd = kd <= 0 ? x->D : kd;
  TSP_CHECK(p.kd <= x->D, "avgpool: window %d larger than D=%d", p.kd, x->D);
  p.OD = x->D - p.kd + 1;
  p.out = out;
  TSP_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "avgpool: out must be 16-byte aligned");
  p.total_warps = static_cast<long long>(x->N) * p.OD * (x->C / 8);
  TSP_CUDA(launch_kernel(avgpool_kernel, dim3(grid_for(p.total_warps * 32, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_l2_normalize_rows(float* x, int32_t rows, int32_t cols, float eps, void* stream) {
  TSP_CHECK(x && rows >= 1 && cols >= 1, "l2_normalize_rows: bad arguments");
  NormP p;
  p.x = x; p.rows = rows; p.cols = cols; p.eps = eps;
  TSP_CUDA(launch_kernel(l2_normalize_rows_kernel, dim3(rows), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_mgfn_rows(const float* feats, int32_t T, int32_t ncrops, int32_t F, const int32_t* bounds, int32_t seg,
                                 int32_t train, float* out, void* stream) {
  TSP_CHECK(feats && out && T >= 1 && ncrops >= 1 && F >= 1, "mgfn_rows: bad arguments (T=%d ncrops=%d F=%d)", T, ncrops, F);
  TSP_CHECK(!train || (bounds != nullptr && seg >= 1), "mgfn_rows: train mode needs the segment boundaries");
  MgfnP p;
  p.feats = feats; p.bounds = bounds; p.out = out; p.T = T; p.ncrops = ncrops; p.F = F; p.seg = seg; p.train = train;
  const int rows = train ? ncrops * seg : T * ncrops;
  TSP_CUDA(launch_kernel(mgfn_rows_kernel, dim3(rows), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_preprocess(const uint8_t* frames, int32_t F, int32_t Hs, int32_t Ws, const int32_t* desc,
                                  int32_t n_out, int32_t crop_h, int32_t crop_w, const tedspad_tensor* y,
                                  int32_t resample, float* frames_f32, void* stream) {
  TSP_CHECK(frames && desc && y, "preprocess: null argument");
  if (check_tensor(*y, "preprocess.y", 1)) return 1;
  TSP_CHECK(n_out >= 1 && y->N == n_out && y->D == 1 && y->C >= 3, "preprocess: y must be [n_out,1,Ho,Wo,>=3]");
  TSP_CHECK(y->H <= PP_MAXOUT && y->W <= PP_MAXOUT, "preprocess: output larger than %d", PP_MAXOUT);
  TSP_CHECK(resample == TEDSPAD_RESAMPLE_AA_FLOAT || resample == TEDSPAD_RESAMPLE_PIL_U8, "preprocess: bad resample");
  TSP_CHECK(crop_h >= 1 && crop_w >= 1 && crop_h <= Hs && crop_w <= Ws && F >= 1, "preprocess: bad crop %dx%d of %dx%d", crop_h,
            crop_w, Hs, Ws);
  TSP_CHECK(y->C <= 8 || (y->C % 8 == 0), "preprocess: y.C=%d", y->C);
  const double sy = static_cast<double>(crop_h) / y->H, sx = static_cast<double>(crop_w) / y->W;
  // taps per axis: hi - lo <= 2 * support + 1
  const int tx = static_cast<int>(ceil(2.0 * (sx < 1 ? 1 : sx))) + 1, ty = static_cast<int>(ceil(2.0 * (sy < 1 ? 1 : sy))) + 1;
  TSP_CHECK(tx <= PP_KMAX + 1 && ty <= PP_KMAX + 1, "preprocess: down-scale factor too large for the %d-tap table", PP_KMAX);
  const bool pil = resample == TEDSPAD_RESAMPLE_PIL_U8;
  // exact maximum tap count per axis (the same arithmetic as axis_entry): the kernels unroll that many taps
  auto max_taps = [&](int in, int out) {
    int m = 1;
    for (int i = 0; i < out; ++i) {
      int lo, hi;
      if (pil) {
        const double scale = static_cast<double>(in) / out, support = scale < 1.0 ? 1.0 : scale, center = (i + 0.5) * scale;
        lo = static_cast<int>(center - support + 0.5); hi = static_cast<int>(center + support + 0.5);
      } else {
        const float scale = static_cast<float>(in) / static_cast<float>(out), support = scale >= 1.f ? scale : 1.f;
        const float center = scale * (i + 0.5f);
        lo = static_cast<int>(center - support + 0.5f); hi = static_cast<int>(center + support + 0.5f);
      }
      lo = lo < 0 ? 0 : lo; hi = hi > in ? in : hi;
      m = std::max(m, hi - lo);
    }
    return m;
  };
  const int mx = max_taps(crop_w, y->W), my = max_taps(crop_h, y->H);
  TSP_CHECK(mx <= PP_KMAX && my <= PP_KMAX, "preprocess: %d x %d taps exceed the %d-tap kernels", mx, my, PP_KMAX);
  const int KXs = mx <= 3 ? 3 : (mx <= 5 ? 5 : 8), KYs = my <= 2 ? 2 : (my <= 5 ? 5 : 8);
  PrepP p;
  p.frames = frames; p.frames_bytes = static_cast<long long>(F) * Hs * Ws * 3;
  p.F = F; p.Hs = Hs; p.Ws = Ws; p.desc = desc; p.n_out = n_out;
  p.crop_h = crop_h; p.crop_w = crop_w; p.resample = resample;
  p.y = make_view(*y);
  p.frames_f32 = frames_f32;
  p.sx = static_cast<float>(crop_w) / static_cast<float>(y->W);
  p.supx = p.sx >= 1.f ? p.sx : 1.f;
  p.invx = p.sx >= 1.f ? 1.f / p.sx : 1.f;
  // band height: as many output rows per block as keep the staged rows + the horizontal-pass planes under 48 KB
  const int nb = crop_w * 3, KY = KYs;
  const bool v8 = y->C == 8 && ((y->ld | y->coff) & 7) == 0, v4 = y->C == 4 && ((y->ld | y->coff) & 3) == 0;
  // the streaming instantiations (FAST, aten path, 2 vertical taps) keep the interpolated rows in registers: no planes,
  // a quarter of the shared memory, twice the resident blocks to hide the staging loads behind
  const bool maybe_stream = !pil && KXs == 3 && KYs == 2 && y->H % PP_FAST_BH == 0 && frames_f32 == nullptr && (v8 || v4);
  p.pitch = static_cast<int>(round_up(nb + 32, 16));
  size_t smem = 0;
  for (p.BH = 8; p.BH >= 1; p.BH >>= 1) {
    p.max_rows = static_cast<int>(ceil((p.BH - 1) * sy + 2.0 * (sy < 1 ? 1 : sy) + 2.0));
    if (p.max_rows > crop_h) p.max_rows = crop_h;
    smem = static_cast<size_t>((p.BH + 1 + p.max_rows + KY * p.BH + 3) & ~3) * 4 +
           static_cast<size_t>(p.max_rows) * p.pitch +
           (maybe_stream && p.BH == PP_FAST_BH ? 0 : static_cast<size_t>(p.max_rows) * 3 * y->W * (pil ? 1 : sizeof(float))) + 16;
    if (smem <= 48 * 1024) break;
  }
  TSP_CHECK(p.BH >= 1, "preprocess: a %d-pixel wide crop does not fit in shared memory", crop_w);
  dim3 grid((y->H + p.BH - 1) / p.BH, n_out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define TSP_PREP(PIL_, KX_, KY_, FAST_) TSP_CUDA(launch_kernel(preprocess_kernel<PIL_, KX_, KY_, FAST_>, grid, dim3(PP_THREADS), smem, st, p))
#define TSP_PREP_Y(PIL_, KX_) do { if (KYs == 2) TSP_PREP(PIL_, KX_, 2, 0); else if (KYs == 5) TSP_PREP(PIL_, KX_, 5, 0); else TSP_PREP(PIL_, KX_, 8, 0); } while (0)
#define TSP_PREP_X(PIL_) do { if (KXs == 3) TSP_PREP_Y(PIL_, 3); else if (KXs == 5) TSP_PREP_Y(PIL_, 5); else TSP_PREP_Y(PIL_, 8); } while (0)
  // the hot configurations (UCF-Crime / XD 240x320 -> 224: 3 x 2 taps; ShanghaiTech 384 -> 224: 5 x 5 taps) in full
  // bands of 8 rows with vector pixel stores and no fp32 copy run the FAST instantiations
  const bool fast = p.BH == PP_FAST_BH && y->H % PP_FAST_BH == 0 && frames_f32 == nullptr && (v8 || v4);
  if (fast && !pil && KXs == 3 && KYs == 2) { if (v8) TSP_PREP(false, 3, 2, 8); else TSP_PREP(false, 3, 2, 4); }
  else if (fast && !pil && KXs == 5 && KYs == 5) { if (v8) TSP_PREP(false, 5, 5, 8); else TSP_PREP(false, 5, 5, 4); }
  else if (fast && pil && KXs == 5 && KYs == 5) { if (v8) TSP_PREP(true, 5, 5, 8); else TSP_PREP(true, 5, 5, 4); }
  else if (pil) TSP_PREP_X(true);
  else TSP_PREP_X(false);
#undef TSP_PREP_X
#undef TSP_PREP_Y
#undef TSP_PREP
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_nchw_to_cl(const float* x, int32_t Cx, const tedspad_tensor* y, void* stream) {
  TSP_CHECK(x && y, "nchw_to_cl: null argument");
  if (check_tensor(*y, "nchw_to_cl.y", y->C == 4 ? 4 : 8)) return 1;
  TSP_CHECK((y->C % 8 == 0 || y->C == 4) && Cx >= 1 && Cx <= y->C, "nchw_to_cl: Cx=%d vs y.C=%d", Cx, y->C);
  CvtP p;
  p.x = x; p.Cx = Cx; p.y = make_view(*y);
  p.total = static_cast<long long>(y->N) * y->D * y->H * y->W * (y->C == 4 ? 1 : y->C / 8);
  TSP_CUDA(launch_kernel(nchw_to_cl_kernel, dim3(grid_for(p.total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tedspad_planes_to_clip(const void* planes, const tedspad_tensor* y, int32_t T, void* stream) {
  TSP_CHECK(planes && y, "planes_to_clip: null argument");
  if (check_tensor(*y, "planes_to_clip.y", 4)) return 1;
  TSP_CHECK((y->C == 4 || y->C == 8) && y->ld % y->C == 0 && y->coff % y->C == 0 && y->W % 8 == 0 && y->D == T && T >= 1,
            "planes_to_clip: y must be [B,T,H,W,4|8] with W %% 8 == 0 (got [%d,%d,%d,%d,%d], T=%d)", y->N, y->D, y->H,
            y->W, y->C, T);
  TSP_CHECK((reinterpret_cast<uintptr_t>(planes) & 15) == 0, "planes_to_clip: planes must be 16-byte aligned");
  P2CP p;
  p.planes = reinterpret_cast<const __nv_bfloat16*>(planes);
  p.y = make_view(*y);
  p.T = T;
  p.total = static_cast<long long>(y->N) * T * y->H * (y->W / 8);
  TSP_CUDA(launch_kernel(planes_to_clip_kernel, dim3(grid_for(p.total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  TSP_CUDA(cudaGetLastError());
  return 0;
}
