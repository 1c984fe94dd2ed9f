// HBM write-bandwidth probe (sm_100a): which store flavour reaches the highest write-only rate?  The stem, the
// up-sampling kernels and preprocessing are judged against the "write ceiling" (DESIGN.md 2): this probe checks whether
// that ceiling (3.92 TB/s from a memset) is a property of the board or of the store instruction used.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tests/probes/hbm_write_probe tests/probes/hbm_write_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// mode 0: st.global.v4 default; 1: st.global.v8 (32 B); 2: st.global.cs.v4; 3: st.global.wt.v4;
// 4: v8 with an L2::evict_first policy; 5: v4 with .L1::no_allocate
template <int MODE>
__global__ void write_linear(uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint32_t v = threadIdx.x;
  if (MODE == 1 || MODE == 4) {
    uint64_t pol = 0;
    if (MODE == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t n32 = n16 / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
      char* p = reinterpret_cast<char*>(dst) + i * 32;
      if (MODE == 1)
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      else
        asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1}, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      uint4* p = dst + i;
      if (MODE == 0) asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 2) asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 3) asm volatile("st.global.wt.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 5) asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    }
  }
}

// The epilogue's pattern: a lane owns one 128-byte pixel row and writes it as 2 x (2 x 32 B) in two passes (chunks),
// a warp covers 32 consecutive rows.  PASSES = 1: the whole 128 B row at once (4 x 32 B back to back).
template <int PASSES>
__global__ void write_rows(char* __restrict__ dst, size_t rows) {
  const uint32_t v = threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    char* p = dst + r * 128;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + c * 64), "r"(v) : "memory");
      asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + c * 64 + 32), "r"(v) : "memory");
      if (PASSES == 2 && c == 0) __nanosleep(200);   // the second chunk follows ~one chunk of arithmetic later
    }
  }
}

// TMA bulk store: each CTA streams 16 KB blocks shared -> global with cp.async.bulk (the CUTLASS epilogue path)
__global__ void write_bulk(char* __restrict__ dst, size_t nblk, int inflight) {
  extern __shared__ __align__(128) char sm[];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    for (size_t b = blockIdx.x; b < nblk; b += gridDim.x) {
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 16384;" ::"l"(dst + b * 16384), "r"(s) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++k >= inflight) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <class F>
static double time_it(F f, size_t bytes, int reps = 6) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  double best = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double tbs = bytes / (ms * 1e-3) / 1e12;
    if (tbs > best) best = tbs;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const size_t bytes = (size_t)4 << 30;   // 4 GiB: 32x the L2
  char* d; CK(cudaMalloc(&d, bytes));
  char* src; CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t n16 = bytes / 16;
  printf("SMs %d, buffer %zu MiB\n", sms, bytes >> 20);
  printf("cudaMemsetAsync                       %.2f TB/s\n", time_it([&] { CK(cudaMemsetAsync(d, 0, bytes)); }, bytes));
  printf("cudaMemcpyAsync d2d (read + write)    %.2f TB/s\n", time_it([&] { CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice)); }, 2 * bytes));
  for (int bpsm : {4, 8, 16}) {
    const int grid = sms * bpsm, blk = 256;
    printf("-- grid %d x %d threads\n", grid, blk);
    printf("st.global.v4 (16 B / thread)          %.2f TB/s\n", time_it([&] { write_linear<0><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("st.global.v8 (32 B / thread)          %.2f TB/s\n", time_it([&] { write_linear<1><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("st.global.cs.v4                       %.2f TB/s\n", time_it([&] { write_linear<2><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("st.global.wt.v4                       %.2f TB/s\n", time_it([&] { write_linear<3><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("st.global.v8 L2::evict_first          %.2f TB/s\n", time_it([&] { write_linear<4><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("st.global.L1::no_allocate.v4          %.2f TB/s\n", time_it([&] { write_linear<5><<<grid, blk>>>((uint4*)d, n16); }, bytes));
    printf("rows: lane = 128 B row, 4 x v8        %.2f TB/s\n", time_it([&] { write_rows<1><<<grid, blk>>>(d, bytes / 128); }, bytes));
    printf("rows: two 64 B chunks 200 ns apart    %.2f TB/s\n", time_it([&] { write_rows<2><<<grid, blk>>>(d, bytes / 128); }, bytes));
  }
  CK(cudaFuncSetAttribute(write_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  for (int bpsm : {1, 2, 4})
    for (int inflight : {4, 8})
      printf("cp.async.bulk s2g 16 KB, %d CTA/SM, %d in flight  %.2f TB/s\n", bpsm, inflight,
             time_it([&] { write_bulk<<<sms * bpsm, 128, 16384>>>(d, bytes / 16384, inflight); }, bytes));
  return 0;
}
