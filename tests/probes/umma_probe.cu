// Hardware probe (test infrastructure, not product code): pins down the tcgen05 behaviours the conv
// engine's feeds rely on, on the actual B200, before any kernel is built on top of them.
//
//   correctness  - SW128 K-major operand descriptors whose start address is shifted by whole 128-byte
//                  rows inside a slab (the kx taps of a 3x3 convolution read ONE slab at +0/+1/+2 rows);
//                - no-swizzle K-major descriptors whose K-adjacent core matrices OVERLAP (LBO = 16 B):
//                  A[r][k] = X[8 r + k], i.e. im2col windows of a stride-s convolution read in place.
//   throughput   - clk per tcgen05.mma (M=128, K=16) for N = 64/128/192/256 with both operands in smem,
//                  all SMs busy: the shared-memory operand-read bound of small-N tiles;
//                - clk per tcgen05.ld 32x32b.x16 chunk for 4 epilogue warps.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#include "../../ted-spad_b200/csrc/ptx.cuh"

using namespace tsp;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off,
                                              uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

struct Cfg {
  int n;                 // UMMA N
  int nk;                // number of K=16 MMAs
  int a_off, a_lbo, a_sbo, a_base, a_layout, a_kstep;
  int b_off, b_lbo, b_sbo, b_base, b_layout, b_kstep;
  int img_bytes;
};

// One CTA: smem image copied verbatim from global, nk MMAs, D (128 x n fp32) written to global.
__global__ void __launch_bounds__(128, 1) probe_mma(const uint8_t* img, Cfg c, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < c.img_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n);
    const uint32_t base = smem_u32(smem);
    for (int k = 0; k < c.nk; ++k) {
      const uint32_t a_addr = base + c.a_off + k * c.a_kstep;
      const uint32_t b_addr = base + c.b_off + k * c.b_kstep;
      const uint32_t abo = c.a_base < 0 ? ((a_addr >> 7) & 7u) : c.a_base;
      const uint32_t bbo = c.b_base < 0 ? ((b_addr >> 7) & 7u) : c.b_base;
      umma_bf16(tmem, make_desc(a_addr, c.a_lbo, c.a_sbo, abo, c.a_layout),
                make_desc(b_addr, c.b_lbo, c.b_sbo, bbo, c.b_layout), idesc, k ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int col = 0; col < c.n; col += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + col, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[row * c.n + col + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Throughput: every CTA issues `reps` x 4 MMAs back to back on fixed smem contents.
__global__ void __launch_bounds__(128, 1) probe_rate(Cfg c, int reps, long long* clk_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n);
    const uint32_t base = smem_u32(smem);
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      ad[k] = make_desc(base + c.a_off + k * c.a_kstep, c.a_lbo, c.a_sbo, 0, c.a_layout);
      bd[k] = make_desc(base + c.b_off + k * c.b_kstep, c.b_lbo, c.b_sbo, 0, c.b_layout);
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad[k], bd[k], idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    clk_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// tcgen05.ld rate: 4 warps, each reads its lane quarter, `cols` columns per pass, `reps` passes.
__global__ void __launch_bounds__(128, 1) probe_ld(int cols, int reps, long long* clk_out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int col = 0; col < cols; col += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += __uint_as_float(v[i]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ host
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static float rnd() { return static_cast<float>((rand() % 17) - 8) / 8.f; }  // exactly representable in bf16

struct Case {
  const char* name;
  Cfg cfg;
  std::vector<uint8_t> img;
  std::vector<float> expect;  // 128 x n
};

// SW128 K-major placement of element (row r, k) of a tile whose row 0 sits at byte `base` (1024-aligned)
static size_t sw128_off(size_t base, int r, int k) {
  const int chunk = k >> 3, within = k & 7;
  return base + (r >> 3) * 1024 + (r & 7) * 128 + (((chunk ^ (r & 7)) & 7) << 4) + within * 2;
}

static void put(std::vector<uint8_t>& img, size_t off, float v) {
  const uint16_t h = f2bf(v);
  memcpy(&img[off], &h, 2);
}

// SW128: A slab of `slab_rows` rows x 64 K at byte 0, B (n x 64) at b_off; MMA reads rows [shift, shift+128)
static Case make_sw128_case(const char* name, int n, int shift, int base_mode) {
  Case c;
  c.name = name;
  const int slab_rows = 144;
  const int b_off = slab_rows * 128;  // 18432, multiple of 1024
  c.img.assign(b_off + n * 128, 0);
  std::vector<float> A(slab_rows * 64), B(n * 64);
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (int r = 0; r < slab_rows; ++r)
    for (int k = 0; k < 64; ++k) put(c.img, sw128_off(0, r, k), A[r * 64 + k]);
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < 64; ++k) put(c.img, sw128_off(b_off, r, k), B[r * 64 + k]);
  c.expect.assign(128 * n, 0.f);
  for (int m = 0; m < 128; ++m)
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
      for (int k = 0; k < 64; ++k) s += A[(m + shift) * 64 + k] * B[j * 64 + k];
      c.expect[m * n + j] = s;
    }
  Cfg& g = c.cfg;
  g.n = n; g.nk = 4;
  g.a_off = shift * 128; g.a_lbo = 16; g.a_sbo = 1024; g.a_base = base_mode; g.a_layout = 2; g.a_kstep = 32;
  g.b_off = b_off; g.b_lbo = 16; g.b_sbo = 1024; g.b_base = 0; g.b_layout = 2; g.b_kstep = 32;
  g.img_bytes = static_cast<int>(c.img.size());
  return c;
}

// no swizzle.  A[r][k] = X[win*r + k] when overlap (win elements per row step, 8 = 16 bytes), else a dense
// chunk-major tile.  field_swap exchanges the LBO/SBO fields to find out which one the hardware calls which.
static Case make_noswz_case(const char* name, int n, int nk, bool overlap, bool field_swap) {
  Case c;
  c.name = name;
  const int K = 16 * nk;
  std::vector<float> A(128 * K), B(n * K);
  size_t a_bytes;
  int a_k_stride, a_m_stride, a_kstep;
  if (overlap) {
    std::vector<float> X(8 * 128 + K + 64);
    for (auto& v : X) v = rnd();
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < K; ++k) A[r * K + k] = X[8 * r + k];
    a_bytes = (X.size() * 2 + 1023) / 1024 * 1024;
    c.img.assign(a_bytes, 0);
    for (size_t i = 0; i < X.size(); ++i) put(c.img, i * 2, X[i]);
    a_k_stride = 16;   // K-adjacent core matrix 16 bytes further: overlapping windows
    a_m_stride = 128;  // next 8-row group
    a_kstep = 32;      // next K=16 slice
  } else {
    for (auto& v : A) v = rnd();
    // chunk-major: chunk c (8 k) of row r at c*2048 + r*16
    a_bytes = static_cast<size_t>(K / 8) * 2048;
    c.img.assign(a_bytes, 0);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < K; ++k) put(c.img, (k >> 3) * 2048 + r * 16 + (k & 7) * 2, A[r * K + k]);
    a_k_stride = 2048;
    a_m_stride = 128;
    a_kstep = 4096;
  }
  for (auto& v : B) v = rnd();
  const size_t b_off = a_bytes;
  c.img.resize(b_off + static_cast<size_t>(K / 8) * n * 16, 0);
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < K; ++k) put(c.img, b_off + (k >> 3) * n * 16 + r * 16 + (k & 7) * 2, B[r * K + k]);
  c.expect.assign(128 * n, 0.f);
  for (int m = 0; m < 128; ++m)
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[j * K + k];
      c.expect[m * n + j] = s;
    }
  Cfg& g = c.cfg;
  g.n = n; g.nk = nk;
  g.a_off = 0; g.a_base = 0; g.a_layout = 0; g.a_kstep = a_kstep;
  g.b_off = static_cast<int>(b_off); g.b_base = 0; g.b_layout = 0; g.b_kstep = 2 * n * 16;
  const int b_k_stride = n * 16, b_m_stride = 128;
  if (!field_swap) {
    g.a_lbo = a_k_stride; g.a_sbo = a_m_stride; g.b_lbo = b_k_stride; g.b_sbo = b_m_stride;
  } else {
    g.a_lbo = a_m_stride; g.a_sbo = a_k_stride; g.b_lbo = b_m_stride; g.b_sbo = b_k_stride;
  }
  g.img_bytes = static_cast<int>((c.img.size() + 15) / 16 * 16);
  c.img.resize(g.img_bytes, 0);
  return c;
}

static bool run_case(Case& c) {
  uint8_t* d_img;
  float* d_out;
  CK(cudaMalloc(&d_img, c.img.size()));
  CK(cudaMalloc(&d_out, 128 * c.cfg.n * 4));
  CK(cudaMemcpy(d_img, c.img.data(), c.img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, 128 * c.cfg.n * 4));
  const int smem = c.cfg.img_bytes + 2048;
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_mma<<<1, 128, smem>>>(d_img, c.cfg, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[CASE] %-44s LAUNCH ERROR %s\n", c.name, cudaGetErrorString(e));
    exit(3);  // context is poisoned: the driver script reruns the remaining cases in a new process
  }
  std::vector<float> out(128 * c.cfg.n);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    const double d = fabs(static_cast<double>(out[i]) - c.expect[i]);
    if (!(d <= 1e-3)) ++bad;
    if (d > maxerr || d != d) maxerr = d;
  }
  printf("[CASE] %-44s %s  bad=%d/%zu maxerr=%g\n", c.name, bad == 0 ? "PASS" : "FAIL", bad, out.size(), maxerr);
  cudaFree(d_img);
  cudaFree(d_out);
  return bad == 0;
}

static void run_rate(const char* name, Cfg c, int sms) {
  long long* d_clk;
  CK(cudaMalloc(&d_clk, sms * 8));
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int reps = 2048;
  for (int it = 0; it < 2; ++it) probe_rate<<<sms, 128, 100 * 1024>>>(c, reps, d_clk);
  CK(cudaDeviceSynchronize());
  std::vector<long long> clk(sms);
  CK(cudaMemcpy(clk.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
  double mean = 0;
  long long mx = 0;
  for (auto v : clk) { mean += v; if (v > mx) mx = v; }
  mean /= sms;
  const double per = mean / (reps * 4.0);
  const double ideal = 128.0 * c.n / 256.0 / 1.0 * (16.0 / 16.0) / 1.0;  // 128*N*16*2 flops / 8192 flop/clk
  printf("[RATE] %-40s N=%3d  clk/mma mean %.1f (max-CTA %.1f)  ideal %.0f  -> %.0f%% of tensor peak\n", name, c.n, per,
         mx / (reps * 4.0), ideal / 1.0, 100.0 * ideal / per);
  cudaFree(d_clk);
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "all";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, cc %d.%d, mode %s\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor, which);
  srand(1234);
  auto want = [&](const char* g) { return !strcmp(which, "all") || !strcmp(which, g); };
  if (want("sw128")) {
    Case c0 = make_sw128_case("sw128 control shift=0 base=0", 64, 0, 0);
    run_case(c0);
    for (int shift : {1, 2, 5, 7, 8, 9}) {
      char* nm = new char[64];
      snprintf(nm, 64, "sw128 shift=%d base_offset=0", shift);
      Case c = make_sw128_case(nm, 64, shift, 0);
      run_case(c);
      nm = new char[64];
      snprintf(nm, 64, "sw128 shift=%d base_offset=(addr>>7)&7", shift);
      Case d = make_sw128_case(nm, 64, shift, -1);
      run_case(d);
    }
    Case c1 = make_sw128_case("sw128 N=192 shift=1 base_offset=0", 192, 1, 0);
    run_case(c1);
  }
  if (want("noswz")) {
    Case a = make_noswz_case("noswz dense   LBO=k-stride SBO=m-stride", 64, 1, false, false);
    run_case(a);
    Case b = make_noswz_case("noswz dense   fields swapped", 64, 1, false, true);
    run_case(b);
    Case a2 = make_noswz_case("noswz dense nk=2 LBO=k SBO=m", 64, 2, false, false);
    run_case(a2);
  }
  if (want("overlap")) {
    Case a = make_noswz_case("noswz OVERLAP LBO=16 SBO=128 nk=1", 64, 1, true, false);
    run_case(a);
    Case a2 = make_noswz_case("noswz OVERLAP LBO=16 SBO=128 nk=2", 64, 2, true, false);
    run_case(a2);
    Case a4 = make_noswz_case("noswz OVERLAP LBO=16 SBO=128 nk=4", 64, 4, true, false);
    run_case(a4);
  }
  if (want("overlap_swapped")) {
    Case b = make_noswz_case("noswz OVERLAP fields swapped nk=1", 64, 1, true, true);
    run_case(b);
    Case b2 = make_noswz_case("noswz OVERLAP fields swapped nk=2", 64, 2, true, true);
    run_case(b2);
  }
  if (want("rate")) {
    const int sms = prop.multiProcessorCount;
    for (int n : {64, 128, 192, 256}) {
      Cfg g;
      memset(&g, 0, sizeof(g));
      g.n = n;
      g.a_off = 0; g.a_lbo = 16; g.a_sbo = 1024; g.a_layout = 2; g.a_kstep = 32;
      g.b_off = 32768; g.b_lbo = 16; g.b_sbo = 1024; g.b_layout = 2; g.b_kstep = 32;
      run_rate("sw128 SS", g, sms);
    }
    {
      Cfg g;
      memset(&g, 0, sizeof(g));
      g.n = 64;
      g.a_off = 0; g.a_lbo = 16; g.a_sbo = 128; g.a_layout = 0; g.a_kstep = 32;
      g.b_off = 32768; g.b_lbo = 1024; g.b_sbo = 128; g.b_layout = 0; g.b_kstep = 2048;
      run_rate("noswz overlapped A (LBO=16)", g, sms);
      g.a_lbo = 2048; g.a_kstep = 4096;
      run_rate("noswz dense A", g, sms);
    }
    long long* d_clk;
    float* d_sink;
    CK(cudaMalloc(&d_clk, sms * 8));
    CK(cudaMalloc(&d_sink, 4));
    for (int cols : {64, 192, 256}) {
      const int reps = 512;
      probe_ld<<<sms, 128>>>(cols, reps, d_clk, d_sink);
      CK(cudaDeviceSynchronize());
      std::vector<long long> clk(sms);
      CK(cudaMemcpy(clk.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
      double mean = 0;
      for (auto v : clk) mean += v;
      mean /= sms;
      printf("[LD]   tcgen05.ld 32x32b.x16 + wait + 16 adds: %d cols x 128 lanes per pass: %.1f clk/pass (%.1f clk per 16-col chunk)\n",
             cols, mean / reps, mean / reps / (cols / 16));
    }
  }
  return 0;
}
