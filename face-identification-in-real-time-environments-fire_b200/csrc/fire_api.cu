// fire_api.cu - library-level pieces of the C ABI (include/fire_b200.h): error reporting,
// device checks, TMA descriptor construction, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "fire_internal.h"

namespace fire {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_f16_2d_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                        uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15) || box_rows == 0 || box_rows > 256)
    return fail(FIRE_ERR_ARG, "tensor map: base/stride must be 16-byte aligned, box rows in [1,256]");
  CUtensorMapSwizzle sw;
  switch (box_cols) {
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: return fail(FIRE_ERR_ARG, "tensor map: box of %u fp16 columns is not a swizzle width (16/32/64)", box_cols);
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return FIRE_OK;
}

int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                     uint32_t box_rows) {
  return make_tmap_f16_2d_ex(out, base, rows, cols, row_stride_bytes, 64, box_rows);
}


int make_tmap_f16_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld_elems,
                       uint64_t w_pitch, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems & 7) || box_w == 0 || box_w > 256 || box_h == 0 || box_h > 256 || box_n == 0 || box_n > 256)
    return fail(FIRE_ERR_ARG, "NHWC tensor map: base/stride must be 16-byte aligned, box extents in [1,256]");
  CUtensorMapSwizzle sw;
  switch (box_c) {
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: return fail(FIRE_ERR_ARG, "NHWC tensor map: box of %u channels is not a swizzle width (16/32/64)", box_c);
  }
  cuuint64_t gdim[4] = {C, W, H, N};
  cuuint64_t gstride[3] = {ld_elems * 2, w_pitch * ld_elems * 2, H * w_pitch * ld_elems * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
  return FIRE_OK;
}


int make_tmap_f16_pos3d(CUtensorMap* out, const void* base, uint64_t C, uint64_t positions, uint64_t N, uint64_t ld_elems,
                        uint32_t box_c, uint32_t box_pos) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems & 7) || box_pos == 0 || box_pos > 256)
    return fail(FIRE_ERR_ARG, "position tensor map: base/stride must be 16-byte aligned, box rows in [1,256]");
  CUtensorMapSwizzle sw;
  switch (box_c) {
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: return fail(FIRE_ERR_ARG, "position tensor map: box of %u channels is not a swizzle width (16/32/64)", box_c);
  }
  cuuint64_t gdim[3] = {C, positions, N};
  cuuint64_t gstride[2] = {ld_elems * 2, positions * ld_elems * 2};
  cuuint32_t box[3] = {box_c, box_pos, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  return FIRE_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_f16_im2col(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld_elems,
                         int kw, int kh, int pad_w, int pad_h, int stride) {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeIm2col not available from the driver");
    fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems & 7) || C % 64)
    return fail(FIRE_ERR_ARG, "im2col tensor map: base/stride must be 16-byte aligned, channels a multiple of 64");
  cuuint64_t gdim[4] = {C, W, H, N};
  cuuint64_t gstride[3] = {ld_elems * 2, W * ld_elems * 2, H * W * ld_elems * 2};
  // base pixels run over [-pad, dim - 1 + pad - (k - 1)] in each spatial dimension (dilation 1); corner arrays are {W, H}
  // like every other array of the tensor map (verified on the 1x3 / 3x1 layers of Block8)
  int lower[2] = {-pad_w, -pad_h};
  int upper[2] = {pad_w - (kw - 1), pad_h - (kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, lower, upper, 64, 128, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FIRE_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
  return FIRE_OK;
}

int device_sm_count() {
  static int n[FIRE_MAX_DEVICES] = {};                 // cached per device: handles may live on different GPUs of one process
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FIRE_MAX_DEVICES) return 148;
  if (n[dev]) return n[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  n[dev] = v;
  return v;
}

int use_device(int device) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return FIRE_OK;
  FIRE_CUDA(cudaSetDevice(device));
  return FIRE_OK;
}

}  // namespace fire

extern "C" {

int fire_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fire::fail(FIRE_ERR_CUDA, "no CUDA device: %s (fire_b200 has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fire::fail(FIRE_ERR_ARG, "device %d out of range [0,%d)", device, n);
  FIRE_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  FIRE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  FIRE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return fire::fail(FIRE_ERR_UNSUPPORTED, "device %d is sm_%d%d; libfire_b200 is built for sm_100a only", device, major,
                      minor);
  return FIRE_OK;
}

const char* fire_last_error(void) { return fire::g_err; }
const char* fire_version(void) { return "fire_b200 0.1.0 (sm_100a)"; }
uint64_t fire_launch_count(void) { return fire::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
