// Host-side helpers shared by the C-ABI translation units: thread-local error text, checks,
// and the lazily resolved driver entry point used to encode TMA tensor maps (resolved through
// cudart so that the library has no link-time dependency on libcuda and loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tedspad.h"

namespace tsp {

void set_error(const char* fmt, ...);

#define TSP_CHECK(cond, ...)        \
  do {                              \
    if (!(cond)) {                  \
      ::tsp::set_error(__VA_ARGS__); \
      return 1;                     \
    }                               \
  } while (0)

#define TSP_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::tsp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

// 2-D bf16 tensor map: dims {inner, outer}, row pitch in bytes, box {box_inner, box_outer},
// 128-byte swizzle, zero fill out of bounds.  Returns 0 on success.
int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer);

// 5-D bf16 tensor map (dims/box innermost first, strides of dims 1..4 in bytes), zero fill out of
// bounds, optional 128-byte swizzle (box[0] * 2 must then be 128 bytes).
int encode_tmap_5d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                        const uint32_t box[5], bool swizzle128, const uint32_t* elem_strides = nullptr);

int num_sms();   // of the CURRENT device (cached per device)

// Per-device one-time setup (cudaFuncSetAttribute is a per-device property of a kernel: a process that touches a
// second GPU must opt every kernel in to > 48 KB of dynamic shared memory there as well).  Returns true exactly once
// per (current device, slot); the caller then performs the setup and, on failure, calls device_once_reset.
enum { ONCE_IGEMM_ATTR = 0, ONCE_SLAB_ATTR = 1, ONCE_SLOTS = 4 };
bool device_once(int slot);
void device_once_reset(int slot);

// Programmatic dependent launch (TEDSPAD_PDL=0 switches it off): every hot-path kernel is launched with
// programmaticStreamSerialization, calls griddepcontrol.launch_dependents when it starts and griddepcontrol.wait
// before its first access to memory a previous kernel may still be using.  The next kernel's CTAs are then placed
// on SMs as the current kernel's CTAs drain, and its prologue (barrier init, TMEM allocation, tensor-map prefetch,
// resident-weight load, resampling tables) overlaps the tail of the current one instead of following it.
bool pdl_enabled();

template <typename P>
inline cudaError_t launch_kernel(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const P& params,
                                 int cluster = 1) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, params);
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// byte address of element (n, d, h, w, c=coff) of the LOGICAL view (halo skipped)
inline int64_t tensor_pixels(const tedspad_tensor& t) {
  return (int64_t)t.N * (t.D + 2 * t.pd) * (t.H + 2 * t.ph) * (t.W + 2 * t.pw);
}

int check_tensor(const tedspad_tensor& t, const char* name, int elem_align);

}  // namespace tsp
