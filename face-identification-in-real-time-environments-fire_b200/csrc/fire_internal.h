// fire_internal.h - host-side helpers shared by the translation units of libfire_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/fire_b200.h"

namespace fire {

// Thread-local error message returned by fire_last_error().
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define FIRE_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::fire::fail(FIRE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define FIRE_LAUNCH_CHECK(what)                                                                      \
  do {                                                                                               \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess)                                                                           \
      return ::fire::fail(FIRE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e));    \
  } while (0)

// 2-D row-major fp16 matrix [rows, cols] with `row_stride_bytes`; box = (64 cols, box_rows), 128-byte swizzle.
// Returns 0 or a FIRE_ERR_* code.  cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda and loads on GPU-less hosts.
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                     uint32_t box_rows);

// Same with a box of `box_cols` fp16 columns (64 / 32 / 16 -> 128B / 64B / 32B swizzle); used for loads and stores.
int make_tmap_f16_2d_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                        uint32_t box_cols, uint32_t box_rows);

// 4-D view {C, W, H, N} of an NHWC fp16 buffer whose pixels are `ld_elems` channels apart (C <= ld_elems: a channel
// slice); box = (box_c channels, box_w, box_h, box_n images), swizzle = box_c * 2 bytes.  Out-of-range coordinates
// (negative included) read as zero and are not written: this is how 'same' padding and ragged edges are handled.
// `w_pitch` = pixels between image rows (>= W).
int make_tmap_f16_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld_elems,
                       uint64_t w_pitch, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n = 1);
// 3-D view {C, positions per image, N} of a buffer whose images are `positions` pixels of `ld_elems` channels: the
// pitched position space written by flat-mode strip convs (box = box_c channels x box_pos positions x 1 image).
int make_tmap_f16_pos3d(CUtensorMap* out, const void* base, uint64_t C, uint64_t positions, uint64_t N, uint64_t ld_elems,
                        uint32_t box_c, uint32_t box_pos);

// im2col view {C, W, H, N} of an NHWC fp16 buffer for a kh x kw convolution (padding pad_h / pad_w on both sides, stride
// `stride`): 64 channels x 128 output pixels per load, 128-byte swizzle (conv_igemm.cuh, tma_a == 2).
int make_tmap_f16_im2col(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld_elems,
                         int kw, int kh, int pad_w, int pad_h, int stride);

constexpr int FIRE_MAX_DEVICES = 64;
int device_sm_count();          // SM count of the CURRENT device
// Every entry point that takes a handle runs on the handle's device: makes it current if it is not (one handle = one device).
int use_device(int device);

// Launch counter (every kernel launch of this library bumps it; bench.py reports it as gpu_launches).
void count_launch(int n = 1);

}  // namespace fire
