// C-ABI plumbing: error text, tensor checks, SM count, TMA tensor-map encoding.
#include <cstddef>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.h"

namespace tsp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_tensor(const tedspad_tensor& t, const char* name, int elem_align) {
  TSP_CHECK(t.ptr != nullptr, "%s: null pointer", name);
  TSP_CHECK(t.N >= 1 && t.D >= 1 && t.H >= 1 && t.W >= 1 && t.C >= 1, "%s: bad extents [%d,%d,%d,%d,%d]", name, t.N,
            t.D, t.H, t.W, t.C);
  TSP_CHECK(t.pd >= 0 && t.ph >= 0 && t.pw >= 0, "%s: negative halo", name);
  TSP_CHECK(t.coff >= 0 && t.coff + t.C <= t.ld, "%s: channel view [%d,%d) exceeds ld=%d", name, t.coff, t.coff + t.C,
            t.ld);
  TSP_CHECK(t.ld % elem_align == 0 && t.coff % elem_align == 0, "%s: ld=%d / coff=%d must be multiples of %d", name,
            t.ld, t.coff, elem_align);
  return 0;
}

constexpr int MAX_DEVICES = 64;
static std::mutex g_dev_mutex;
static int g_sms[MAX_DEVICES] = {0};
static bool g_once[MAX_DEVICES][ONCE_SLOTS] = {{false}};

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return -1;
  return dev;
}

int num_sms() {
  const int dev = current_device();
  if (dev < 0) return 1;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (g_sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 1;
    g_sms[dev] = n;
  }
  return g_sms[dev];
}

bool device_once(int slot) {
  const int dev = current_device();
  if (dev < 0) return true;   // unknown device: redo the (idempotent) setup every time
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (g_once[dev][slot]) return false;
  g_once[dev][slot] = true;
  return true;
}

void device_once_reset(int slot) {
  const int dev = current_device();
  if (dev < 0) return;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  g_once[dev][slot] = false;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  TSP_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  TSP_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch_bytes % 16 == 0,
            "tensor map: base/pitch must be 16-byte aligned (base=%p pitch=%llu)", base,
            (unsigned long long)pitch_bytes);
  TSP_CHECK(box_inner * 2 == 128 && box_outer >= 1 && box_outer <= 256, "tensor map: bad box %ux%u", box_inner,
            box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu pitch %llu)", (int)r,
            (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
  return 0;
}

int encode_tmap_5d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                        const uint32_t box[5], bool swizzle128, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  TSP_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  TSP_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p must be 16-byte aligned", base);
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    if (elem_strides != nullptr) es[i] = elem_strides[i];   // traversal stride: ceil(box / stride) elements are loaded
    TSP_CHECK(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "tensor map: bad dim/box %d: %llu / %u", i,
              (unsigned long long)dims[i], box[i]);
  }
  for (int i = 0; i < 4; ++i) {
    st[i] = strides_bytes[i];
    TSP_CHECK(st[i] % 16 == 0, "tensor map: stride %d = %llu bytes is not a multiple of 16", i,
              (unsigned long long)st[i]);
  }
  TSP_CHECK((box[0] * 2) % 16 == 0 && (!swizzle128 || box[0] * 2 == 128), "tensor map: bad inner box %u", box[0]);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), d, st, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(5d) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tsp

namespace tsp {
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("TEDSPAD_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace tsp

extern "C" int tedspad_abi_version(void) { return TEDSPAD_ABI_VERSION; }

extern "C" int tedspad_abi_layout(int32_t* out, int32_t n) {
  const int32_t v[10] = {(int32_t)sizeof(tedspad_tensor), (int32_t)sizeof(tedspad_conv), (int32_t)sizeof(tedspad_conv_slab),
                         (int32_t)sizeof(tedspad_slab_plan), (int32_t)offsetof(tedspad_conv, y2),
                         (int32_t)offsetof(tedspad_conv_slab, kind), (int32_t)offsetof(tedspad_conv_slab, res),
                         (int32_t)offsetof(tedspad_conv_slab, oc_clip), (int32_t)offsetof(tedspad_conv_slab, stack_rows),
                         (int32_t)offsetof(tedspad_slab_plan, tab)};
  for (int i = 0; i < 10 && i < n; ++i) out[i] = v[i];
  return 10;
}
extern "C" int tedspad_num_sms(void) { return tsp::num_sms(); }
extern "C" const char* tedspad_last_error(void) { return tsp::g_err; }
