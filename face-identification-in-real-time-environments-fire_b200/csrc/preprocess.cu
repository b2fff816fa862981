// preprocess.cu - fused crop + resize + normalise for batches of face boxes (K1 of DESIGN.md).
//
// Reference behaviour being replaced (per face, on the CPU):
//   modules/face_recognition.py:412-420   x,y,w,h = max(0,.) each; face = image[y:y+h, x:x+w]
//   modules/encoder.py:19-27              cv2.resize(face,(160,160),INTER_AREA) on uint8 -> /255.0
//
// FIRE_PRE_REFERENCE reproduces cv::resize(INTER_AREA) for 8UC3 bit for bit (every branch of
// imgproc/resize.cpp that this call can take: copy, integer-scale box mean, fractional area
// tables in float, fixed-point linear for up-scaled axes) and re-quantises to uint8 like the
// reference does before dividing by 255.  FIRE_PRE_NORTHSTAR is the additive mode named by the
// north star: float half-pixel bilinear + per-crop prewhiten (davidsandberg facenet.prewhiten).
//
// Layout: frames are uint8 HWC3 in HBM; the output is the network's input layout, fp16 space-to-depth
// [80][80][16] (2 x 2 pixel blocks, 12 channels used, pixel scale 0..255; see store_s2d), 204.8 KB per crop.  The kernel is HBM/L2-bound integer/byte work; it
// deliberately stays off the tensor cores.
#include <cfloat>
#include <cmath>

#include "fire_common.cuh"
#include "fire_internal.h"

namespace fire {

constexpr int PRE_OUT = 160;
constexpr int PRE_ROWS_PER_BLOCK = 16;
constexpr int PRE_THREADS = 256;

enum { PM_EMPTY = 0, PM_COPY = 1, PM_FAST = 2, PM_AREA = 3, PM_LINEAR = 4 };

struct AreaEntry {        // source span of one output coordinate, cv::computeResizeAreaTab order
  int s_left;             // index of the left partial cell (valid if has_left)
  int s_mid0, n_mid;      // full cells [s_mid0, s_mid0 + n_mid)
  int s_right;            // right partial cell (valid if has_right)
  float a_left, a_mid, a_right;
  int has_left, has_right;
};
struct LinEntry {
  int ofs;                // source index
  int a0, a1;             // 11-bit fixed-point weights
  int edge;               // x only: dx >= xmax -> single tap * 2048
};

struct CropGeom {
  const uint8_t* base;    // first byte of the crop (row y0, col x0)
  long long stride;
  int cw, ch, mode, iscale_x, iscale_y;
  double scale_x, scale_y, inv_scale_x, inv_scale_y;
};

__device__ __forceinline__ int cv_floor_dev(double v) { int i = static_cast<int>(v); return i - (i > v); }
__device__ __forceinline__ int cv_ceil_dev(double v) { int i = static_cast<int>(v); return i + (i < v); }

// modules/face_recognition.py:412-420 + numpy slice clipping
__device__ __forceinline__ void crop_geometry(const uint8_t* frames, const int64_t* fd, const int32_t* box, CropGeom& g) {
  const long long off = fd[0];
  const int H = static_cast<int>(fd[1]), W = static_cast<int>(fd[2]);
  g.stride = fd[3];
  int x = max(0, box[0]), y = max(0, box[1]), w = max(0, box[2]), h = max(0, box[3]);
  const int x1 = min(W, x + w), y1 = min(H, y + h);
  const int x0 = min(x, W), y0 = min(y, H);
  g.cw = x1 - x0; g.ch = y1 - y0;
  g.base = frames + off + static_cast<long long>(y0) * g.stride + static_cast<long long>(x0) * 3;
  if (g.cw <= 0 || g.ch <= 0) { g.mode = PM_EMPTY; return; }
  if (g.cw == PRE_OUT && g.ch == PRE_OUT) { g.mode = PM_COPY; return; }
  g.inv_scale_x = static_cast<double>(PRE_OUT) / g.cw; g.inv_scale_y = static_cast<double>(PRE_OUT) / g.ch;
  g.scale_x = 1. / g.inv_scale_x; g.scale_y = 1. / g.inv_scale_y;
  g.iscale_x = __double2int_rn(g.scale_x); g.iscale_y = __double2int_rn(g.scale_y);
  if (g.scale_x >= 1 && g.scale_y >= 1) {
    const bool fast = fabs(g.scale_x - g.iscale_x) < DBL_EPSILON && fabs(g.scale_y - g.iscale_y) < DBL_EPSILON;
    g.mode = fast ? PM_FAST : PM_AREA;
  } else {
    g.mode = PM_LINEAR;
  }
}

__device__ __forceinline__ void area_entry(int d, int ssize, double scale, AreaEntry& e) {
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, ssize - fsx1);
  int sx1 = cv_ceil_dev(fsx1), sx2 = cv_floor_dev(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  e.has_left = (sx1 - fsx1 > 1e-3);
  e.s_left = sx1 - 1;
  e.a_left = static_cast<float>((sx1 - fsx1) / cell);
  e.s_mid0 = sx1; e.n_mid = max(0, sx2 - sx1);
  e.a_mid = static_cast<float>(1.0 / cell);
  e.has_right = (fsx2 - sx2 > 1e-3);
  e.s_right = sx2;
  e.a_right = static_cast<float>(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
}

__device__ __forceinline__ void linear_entry(int d, int ssize, double scale, double inv_scale, bool is_x, LinEntry& e) {
  int s = cv_floor_dev(d * scale);
  float f = static_cast<float>((d + 1) - (s + 1) * inv_scale);
  f = f <= 0 ? 0.f : f - static_cast<float>(cv_floor_dev(f));
  e.edge = 0;
  if (is_x) {
    if (s < 0) { f = 0; s = 0; }
    if (s + 1 >= ssize) {
      e.edge = 1;                        // contributes to xmax = min over such dx (monotone in dx)
      if (s >= ssize - 1) { f = 0; s = ssize - 1; }
    }
  }
  e.ofs = s;
  const float c0 = 1.f - f;
  e.a0 = min(32767, __float2int_rn(__fmul_rn(c0, 2048.f)));
  e.a1 = min(32767, __float2int_rn(__fmul_rn(f, 2048.f)));
}

// Network-input layout (space-to-depth, see fire_b200/netplan.py): fp16 [box][80][80][16]; position (Y, X) holds the
// 2 x 2 pixel block (2Y + dy, 2X + dx) as channels (dy * 2 + dx) * 3 + c, channels 12..15 zero.  Lanes 2j / 2j+1 of
// a warp hold horizontally adjacent pixels, so the even lane collects its neighbour's three values with one shuffle
// round and writes 12 contiguous bytes; the even lane of an odd row also writes the 8 bytes of zero padding.
// MUST be called by all 32 lanes of the warp (it shuffles).
__device__ __forceinline__ void store_s2d(float a0, float a1, float a2, int box, int dy, int dx, __half* out_f16) {
  const float b0 = __shfl_xor_sync(0xffffffffu, a0, 1), b1 = __shfl_xor_sync(0xffffffffu, a1, 1), b2 = __shfl_xor_sync(0xffffffffu, a2, 1);
  if ((dx & 1) == 0) {
    uint32_t* q = reinterpret_cast<uint32_t*>(out_f16 + ((static_cast<size_t>(box) * (PRE_OUT / 2) + (dy >> 1)) * (PRE_OUT / 2) + (dx >> 1)) * 16 +
                                              (dy & 1) * 6);
    q[0] = pack_f16x2_sat(a0, a1);
    q[1] = pack_f16x2_sat(a2, b0);
    q[2] = pack_f16x2_sat(b1, b2);
    if (dy & 1) { q[3] = 0u; q[4] = 0u; }             // halfs 12..15 (q points at half 6 here)
  }
}

// One output pixel of cv::resize(INTER_AREA) on 8UC3 (every branch resize.cpp can take for this call), re-quantised to
// uint8 like the reference does before dividing by 255.  `src` / `st` address the crop: pixel (sy, sx) channel c is
// src[sy * st + sx * 3 + c]; they point either at the frame in HBM or at the CTA's staged copy in shared memory.
__device__ __forceinline__ void reference_pixel(int mode, const CropGeom& g, const uint8_t* __restrict__ src, long long st, int dx, int dy,
                                                const AreaEntry& eax, const AreaEntry& eay, const LinEntry& elx, const LinEntry& ely, int xmax,
                                                int (&v)[3]) {
  v[0] = v[1] = v[2] = 0;
  if (mode == PM_COPY) {
    const uint8_t* s = src + dy * st + dx * 3;
    v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
  } else if (mode == PM_FAST) {
    const int ix = g.iscale_x, iy = g.iscale_y;
    int sum[3] = {0, 0, 0};
    for (int yy = 0; yy < iy; ++yy) {
      const uint8_t* s = src + static_cast<long long>(dy * iy + yy) * st + static_cast<long long>(dx) * ix * 3;
      for (int xx = 0; xx < ix; ++xx) { sum[0] += s[xx * 3]; sum[1] += s[xx * 3 + 1]; sum[2] += s[xx * 3 + 2]; }
    }
    if (ix == 2 && iy == 2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (sum[c] + 2) >> 2;
    } else {
      const float scale = 1.f / static_cast<float>(ix * iy);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = min(255, max(0, __float2int_rn(__fmul_rn(static_cast<float>(sum[c]), scale))));
    }
  } else if (mode == PM_AREA) {
    const AreaEntry ex = eax;
    const AreaEntry ey = eay;
    float sum[3] = {0.f, 0.f, 0.f};
    bool first = true;
    const int ny = ey.has_left + ey.n_mid + ey.has_right;
    for (int j = 0; j < ny; ++j) {
      int sy; float beta;
      if (ey.has_left && j == 0) { sy = ey.s_left; beta = ey.a_left; }
      else if (j - ey.has_left < ey.n_mid) { sy = ey.s_mid0 + j - ey.has_left; beta = ey.a_mid; }
      else { sy = ey.s_right; beta = ey.a_right; }
      const uint8_t* s = src + static_cast<long long>(sy) * st;
      float buf[3] = {0.f, 0.f, 0.f};
      if (ex.has_left) {
        const uint8_t* q = s + ex.s_left * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn(static_cast<float>(q[c]), ex.a_left));
      }
      for (int k = 0; k < ex.n_mid; ++k) {
        const uint8_t* q = s + (ex.s_mid0 + k) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn(static_cast<float>(q[c]), ex.a_mid));
      }
      if (ex.has_right) {
        const uint8_t* q = s + ex.s_right * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn(static_cast<float>(q[c]), ex.a_right));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        sum[c] = first ? __fmul_rn(beta, buf[c]) : __fadd_rn(sum[c], __fmul_rn(beta, buf[c]));
      first = false;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = min(255, max(0, __float2int_rn(sum[c])));
  } else if (mode == PM_LINEAR) {
    const LinEntry ex = elx;
    const LinEntry ey = ely;
    const int r0 = min(max(ey.ofs, 0), g.ch - 1), r1 = min(max(ey.ofs + 1, 0), g.ch - 1);
    const uint8_t* s0 = src + static_cast<long long>(r0) * st + ex.ofs * 3;
    const uint8_t* s1 = src + static_cast<long long>(r1) * st + ex.ofs * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int h0, h1;
      if (dx < xmax) {
        h0 = s0[c] * ex.a0 + s0[3 + c] * ex.a1;
        h1 = s1[c] * ex.a0 + s1[3 + c] * ex.a1;
      } else {
        h0 = s0[c] * 2048; h1 = s1[c] * 2048;
      }
      v[c] = ((((ey.a0 * (h0 >> 4)) >> 16) + ((ey.a1 * (h1 >> 4)) >> 16) + 2) >> 2) & 0xFF;
    }
  }
}

// One CTA = one box x 16 output rows.
//   1. thread 0 applies the crop rule; the span tables of the 160 output columns and 16 output rows are built in shared memory;
//   2. the source rows this row block needs are STAGED in shared memory with 16-byte loads (one coalesced uint4 per lane
//      along the row, from the 16-byte-aligned address below the crop's first byte to the one above its last; the few
//      chunks that would cross the frame's own bytes are read bytewise) - every source byte crosses HBM/L2 -> SM once per
//      row block, whatever the tap overlap.  Blocks whose span does not fit (very large boxes) read the frame directly;
//   3. one thread finishes one network-input POSITION (the 2 x 2 pixel block of the space-to-depth layout): 4 pixels ->
//      16 fp16 channels -> one aligned 32-byte store; a warp writes 1 KB of consecutive bytes.
constexpr int PRE_STAGE_BYTES = 32 * 1024;      // + 9 KB of tables: five CTAs per SM (measured round 2: the kernel is occupancy-bound at three)

__global__ void __launch_bounds__(PRE_THREADS)
preprocess_reference_kernel(const uint8_t* __restrict__ frames, const int64_t* __restrict__ frame_desc,
                            const int32_t* __restrict__ boxes, const int32_t* __restrict__ box_frame, int swap_rb,
                            __half* __restrict__ out_f16, float* __restrict__ out_f32, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t stage[];
  __shared__ CropGeom g;
  __shared__ AreaEntry ax[PRE_OUT];
  __shared__ AreaEntry ay[PRE_ROWS_PER_BLOCK];
  __shared__ LinEntry lx[PRE_OUT];
  __shared__ LinEntry ly[PRE_ROWS_PER_BLOCK];
  __shared__ int s_xmax, s_row_lo, s_row_hi;
  __shared__ long long s_frame_lo, s_frame_hi;     // byte range of the frame inside `frames`

  const int box = blockIdx.x;
  const int dy0 = blockIdx.y * PRE_ROWS_PER_BLOCK;
  if (threadIdx.x == 0) {
    const int64_t* fd = frame_desc + 4 * static_cast<long long>(box_frame[box]);
    crop_geometry(frames, fd, boxes + 4 * box, g);
    s_frame_lo = fd[0];
    s_frame_hi = fd[0] + fd[1] * fd[3];
    s_xmax = PRE_OUT;
    if (status && blockIdx.y == 0) status[box] = g.mode == PM_EMPTY ? 1 : 0;
  }
  __syncthreads();
  const int mode = g.mode;
  if (mode == PM_AREA) {
    for (int i = threadIdx.x; i < PRE_OUT + PRE_ROWS_PER_BLOCK; i += PRE_THREADS) {
      if (i < PRE_OUT) area_entry(i, g.cw, g.scale_x, ax[i]);
      else area_entry(dy0 + i - PRE_OUT, g.ch, g.scale_y, ay[i - PRE_OUT]);
    }
  } else if (mode == PM_LINEAR) {
    for (int i = threadIdx.x; i < PRE_OUT + PRE_ROWS_PER_BLOCK; i += PRE_THREADS) {
      if (i < PRE_OUT) {
        linear_entry(i, g.cw, g.scale_x, g.inv_scale_x, true, lx[i]);
        if (lx[i].edge) atomicMin(&s_xmax, i);
      } else {
        linear_entry(dy0 + i - PRE_OUT, g.ch, g.scale_y, g.inv_scale_y, false, ly[i - PRE_OUT]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                       // source rows [lo, hi) this row block reads
    int lo = 0, hi = 0;
    if (mode == PM_COPY) { lo = dy0; hi = dy0 + PRE_ROWS_PER_BLOCK; }
    else if (mode == PM_FAST) { lo = dy0 * g.iscale_y; hi = (dy0 + PRE_ROWS_PER_BLOCK) * g.iscale_y; }
    else if (mode == PM_AREA) {
      lo = g.ch; hi = 0;
      for (int i = 0; i < PRE_ROWS_PER_BLOCK; ++i) {
        const AreaEntry& e = ay[i];
        if (e.has_left) { lo = min(lo, e.s_left); hi = max(hi, e.s_left + 1); }
        if (e.n_mid > 0) { lo = min(lo, e.s_mid0); hi = max(hi, e.s_mid0 + e.n_mid); }
        if (e.has_right) { lo = min(lo, e.s_right); hi = max(hi, e.s_right + 1); }
      }
    } else if (mode == PM_LINEAR) {
      lo = g.ch; hi = 0;
      for (int i = 0; i < PRE_ROWS_PER_BLOCK; ++i) {
        const int r0 = min(max(ly[i].ofs, 0), g.ch - 1), r1 = min(max(ly[i].ofs + 1, 0), g.ch - 1);
        lo = min(lo, r0); hi = max(hi, r1 + 1);
      }
    }
    s_row_lo = max(0, min(lo, g.ch)); s_row_hi = max(s_row_lo, min(hi, g.ch));
  }
  __syncthreads();
  const int xmax = s_xmax;
  const uint8_t* __restrict__ src = g.base;
  long long st = g.stride;

  if (mode != PM_EMPTY) {
    // stage rows [row_lo, row_hi) x bytes [0, cw * 3) of the crop; smem row r holds the 16-byte-aligned span around them
    const int row_lo = s_row_lo, n_rows = s_row_hi - s_row_lo;
    const int lead = static_cast<int>(reinterpret_cast<uintptr_t>(g.base) & 15);      // same for every row iff the stride is a multiple of 16
    const int pitch = (lead + g.cw * 3 + 15) & ~15;
    if ((g.stride & 15) == 0 && n_rows > 0 && static_cast<long long>(n_rows) * pitch <= PRE_STAGE_BYTES) {
      const int chunks = pitch >> 4;
      const uint8_t* row0 = g.base - lead + static_cast<long long>(row_lo) * g.stride;
      const uint8_t* f_lo = frames + s_frame_lo;
      const uint8_t* f_hi = frames + s_frame_hi;
      for (int i = threadIdx.x; i < n_rows * chunks; i += PRE_THREADS) {
        const int r = i / chunks, c = i - r * chunks;
        const uint8_t* gp = row0 + static_cast<long long>(r) * g.stride + c * 16;
        uint4 q;
        if (gp >= f_lo && gp + 16 <= f_hi) {
          q = __ldg(reinterpret_cast<const uint4*>(gp));
        } else {                                 // the chunk hangs over the frame's first / last byte
          uint8_t b[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) b[k] = (gp + k >= f_lo && gp + k < f_hi) ? gp[k] : static_cast<uint8_t>(0);
          q = *reinterpret_cast<uint4*>(b);
        }
        *reinterpret_cast<uint4*>(stage + static_cast<size_t>(r) * pitch + c * 16) = q;
      }
      __syncthreads();
      src = stage + lead - static_cast<long long>(row_lo) * pitch;      // crop-relative addressing into the staged rows
      st = pitch;
    }
  }

  constexpr int POS_W = PRE_OUT / 2;
  for (int t = threadIdx.x; t < (PRE_ROWS_PER_BLOCK / 2) * POS_W; t += PRE_THREADS) {
    const int yl = t / POS_W, X = t - yl * POS_W;
    float h[12];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ry = 2 * yl + a, dx = 2 * X + b;
        int v[3];
        reference_pixel(mode, g, src, st, dx, dy0 + ry, ax[dx], ay[ry], lx[dx], ly[ry], xmax, v);
        if (swap_rb) { const int tmp = v[0]; v[0] = v[2]; v[2] = tmp; }
        h[(a * 2 + b) * 3 + 0] = static_cast<float>(v[0]);
        h[(a * 2 + b) * 3 + 1] = static_cast<float>(v[1]);
        h[(a * 2 + b) * 3 + 2] = static_cast<float>(v[2]);
      }
    const int Y = (dy0 >> 1) + yl;
    if (out_f16) {
      uint4* o = reinterpret_cast<uint4*>(out_f16 + ((static_cast<size_t>(box) * POS_W + Y) * POS_W + X) * 16);
      o[0] = make_uint4(pack_f16x2_sat(h[0], h[1]), pack_f16x2_sat(h[2], h[3]), pack_f16x2_sat(h[4], h[5]), pack_f16x2_sat(h[6], h[7]));
      o[1] = make_uint4(pack_f16x2_sat(h[8], h[9]), pack_f16x2_sat(h[10], h[11]), 0u, 0u);
    }
    if (out_f32) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float2* o = reinterpret_cast<float2*>(out_f32 + ((static_cast<size_t>(box) * PRE_OUT + 2 * Y + a) * PRE_OUT + 2 * X) * 3);
        o[0] = make_float2(__fdiv_rn(h[a * 6 + 0], 255.0f), __fdiv_rn(h[a * 6 + 1], 255.0f));
        o[1] = make_float2(__fdiv_rn(h[a * 6 + 2], 255.0f), __fdiv_rn(h[a * 6 + 3], 255.0f));
        o[2] = make_float2(__fdiv_rn(h[a * 6 + 4], 255.0f), __fdiv_rn(h[a * 6 + 5], 255.0f));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NORTHSTAR mode: half-pixel bilinear in float + per-crop prewhiten.  One block per crop; pass 1
// accumulates sum / sum-of-squares (double), pass 2 recomputes the taps (L1/L2 hits) and writes
// y = (x - mean) / max(std, 1/sqrt(n)).  out_f16 carries 255*y (the engine folds 1/255 into conv 1).
constexpr int NS_THREADS = 1024;

__device__ __forceinline__ void bilinear_px(const uint8_t* __restrict__ src, long long st, int cw, int ch, float sxs,
                                            float sys, int dx, int dy, float (&v)[3]) {
  float fx = (dx + 0.5f) * sxs - 0.5f, fy = (dy + 0.5f) * sys - 0.5f;
  fx = fminf(fmaxf(fx, 0.f), static_cast<float>(cw - 1));
  fy = fminf(fmaxf(fy, 0.f), static_cast<float>(ch - 1));
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const int x1 = min(x0 + 1, cw - 1), y1 = min(y0 + 1, ch - 1);
  const float tx = fx - x0, ty = fy - y0;
  const uint8_t* p00 = src + static_cast<long long>(y0) * st + x0 * 3;
  const uint8_t* p01 = src + static_cast<long long>(y0) * st + x1 * 3;
  const uint8_t* p10 = src + static_cast<long long>(y1) * st + x0 * 3;
  const uint8_t* p11 = src + static_cast<long long>(y1) * st + x1 * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float top = __fadd_rn(__fmul_rn(1.f - tx, static_cast<float>(p00[c])), __fmul_rn(tx, static_cast<float>(p01[c])));
    const float bot = __fadd_rn(__fmul_rn(1.f - tx, static_cast<float>(p10[c])), __fmul_rn(tx, static_cast<float>(p11[c])));
    v[c] = __fadd_rn(__fmul_rn(1.f - ty, top), __fmul_rn(ty, bot));
  }
}

__global__ void __launch_bounds__(NS_THREADS)
preprocess_northstar_kernel(const uint8_t* __restrict__ frames, const int64_t* __restrict__ frame_desc,
                            const int32_t* __restrict__ boxes, const int32_t* __restrict__ box_frame, int swap_rb,
                            __half* __restrict__ out_f16, float* __restrict__ out_f32, int32_t* __restrict__ status) {
  __shared__ CropGeom g;
  __shared__ double red[2][NS_THREADS / 32];
  __shared__ float s_mean, s_inv;
  const int box = blockIdx.x;
  if (threadIdx.x == 0) {
    crop_geometry(frames, frame_desc + 4 * static_cast<long long>(box_frame[box]), boxes + 4 * box, g);
    if (status) status[box] = g.mode == PM_EMPTY ? 1 : 0;
  }
  __syncthreads();
  const size_t pix0 = static_cast<size_t>(box) * PRE_OUT * PRE_OUT;
  if (g.mode == PM_EMPTY) {
    for (int t = threadIdx.x; t < PRE_OUT * PRE_OUT; t += NS_THREADS) {
      if (out_f16 && t < PRE_OUT * PRE_OUT / 2) *reinterpret_cast<uint4*>(out_f16 + pix0 * 4 + static_cast<size_t>(t) * 8) = make_uint4(0, 0, 0, 0);   // 80*80*16 halfs
      if (out_f32) { out_f32[(pix0 + t) * 3] = 0.f; out_f32[(pix0 + t) * 3 + 1] = 0.f; out_f32[(pix0 + t) * 3 + 2] = 0.f; }
    }
    return;
  }
  const float sxs = static_cast<float>(g.cw) / PRE_OUT, sys = static_cast<float>(g.ch) / PRE_OUT;
  double s = 0., ss = 0.;
  for (int t = threadIdx.x; t < PRE_OUT * PRE_OUT; t += NS_THREADS) {
    float v[3];
    bilinear_px(g.base, g.stride, g.cw, g.ch, sxs, sys, t % PRE_OUT, t / PRE_OUT, v);
#pragma unroll
    for (int c = 0; c < 3; ++c) { s += v[c]; ss += static_cast<double>(v[c]) * v[c]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0., b = 0.;
    for (int i = 0; i < NS_THREADS / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    const double n = 3.0 * PRE_OUT * PRE_OUT;
    const double mean = a / n;
    const double var = fmax(b / n - mean * mean, 0.);
    const double sd = fmax(sqrt(var), 1.0 / sqrt(n));
    s_mean = static_cast<float>(mean);
    s_inv = static_cast<float>(1.0 / sd);
  }
  __syncthreads();
  const float mean = s_mean, inv = s_inv;
  for (int t = threadIdx.x; t < PRE_OUT * PRE_OUT; t += NS_THREADS) {
    float v[3];
    bilinear_px(g.base, g.stride, g.cw, g.ch, sxs, sys, t % PRE_OUT, t / PRE_OUT, v);
    float y[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = (v[c] - mean) * inv;
    if (swap_rb) { const float tmp = y[0]; y[0] = y[2]; y[2] = tmp; }
    if (out_f16) store_s2d(y[0] * 255.f, y[1] * 255.f, y[2] * 255.f, box, t / PRE_OUT, t % PRE_OUT, out_f16);
    if (out_f32) { out_f32[(pix0 + t) * 3] = y[0]; out_f32[(pix0 + t) * 3 + 1] = y[1]; out_f32[(pix0 + t) * 3 + 2] = y[2]; }
  }
}


// ------------------------------------------------------------------------------------------------
// Aligned crop of the enrol path: cv2.warpAffine(image, M, (160, 160)) with INTER_LINEAR and a
// constant zero border, exactly as OpenCV computes it (imgwarp.cpp): the forward 2x3 matrix is inverted in double,
// source coordinates are 10-bit fixed point with 5 interpolation bits, the four taps are blended with the 15-bit
// weight table of remapBilinear.  Reference call sites: yunet_face_detector.py:135-160 (and the MediaPipe /
// RetinaFace twins), followed by [:, :, ::-1] (swap_rb).  Double arithmetic uses explicit _rn intrinsics: the host
// code this must match bit for bit is compiled without FMA contraction.
__device__ short g_warp_tab[1024 * 4];          // BilinearTab_i, built on the host by build_warp_tab()

__global__ void __launch_bounds__(PRE_OUT)
align_warp_kernel(const uint8_t* __restrict__ frames, const int64_t* __restrict__ frame_desc, const double* __restrict__ matrices,
                  const int32_t* __restrict__ face_frame, int swap_rb, uint8_t* __restrict__ out_u8, __half* __restrict__ out_f16) {
  __shared__ double sM[6];
  const int face = blockIdx.x, x = threadIdx.x;
  const int64_t* fd = frame_desc + 4 * static_cast<long long>(face_frame[face]);
  const uint8_t* __restrict__ src = frames + fd[0];
  const int sh = static_cast<int>(fd[1]), sw = static_cast<int>(fd[2]);
  const long long sstride = fd[3];
  if (threadIdx.x == 0) {
    const double* Mf = matrices + 6 * static_cast<long long>(face);
    double M0 = Mf[0], M1 = Mf[1], M2 = Mf[2], M3 = Mf[3], M4 = Mf[4], M5 = Mf[5];
    double D = __dsub_rn(__dmul_rn(M0, M4), __dmul_rn(M1, M3));
    D = D != 0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(M4, D), A22 = __dmul_rn(M0, D);
    M0 = A11; M1 = __dmul_rn(M1, -D);
    M3 = __dmul_rn(M3, -D); M4 = A22;
    const double b1 = __dsub_rn(__dmul_rn(-M0, M2), __dmul_rn(M1, M5));
    const double b2 = __dsub_rn(__dmul_rn(-M3, M2), __dmul_rn(M4, M5));
    sM[0] = M0; sM[1] = M1; sM[2] = b1; sM[3] = M3; sM[4] = M4; sM[5] = b2;
  }
  __syncthreads();
  const double M0 = sM[0], M1 = sM[1], M2 = sM[2], M3 = sM[3], M4 = sM[4], M5 = sM[5];
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M0, static_cast<double>(x)), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M3, static_cast<double>(x)), 1024.0));
  const int y_begin = blockIdx.y * PRE_ROWS_PER_BLOCK;
  for (int y = y_begin; y < y_begin + PRE_ROWS_PER_BLOCK; ++y) {
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M1, static_cast<double>(y)), M2), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M4, static_cast<double>(y)), M5), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const short* w = g_warp_tab + ((Y & 31) * 32 + (X & 31)) * 4;
    const int w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    int v[3] = {0, 0, 0};
    if (static_cast<unsigned>(sx) < static_cast<unsigned>(sw - 1) && static_cast<unsigned>(sy) < static_cast<unsigned>(sh - 1)) {
      const uint8_t* s0 = src + static_cast<long long>(sy) * sstride + sx * 3;
      const uint8_t* s1 = s0 + sstride;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (s0[c] * w0 + s0[c + 3] * w1 + s1[c] * w2 + s1[c + 3] * w3 + (1 << 14)) >> 15;
    } else if (!(sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0)) {
      const bool x0ok = sx >= 0 && sx < sw, x1ok = sx + 1 >= 0 && sx + 1 < sw, y0ok = sy >= 0 && sy < sh, y1ok = sy + 1 >= 0 && sy + 1 < sh;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int p0 = (x0ok && y0ok) ? src[static_cast<long long>(sy) * sstride + sx * 3 + c] : 0;
        const int p1 = (x1ok && y0ok) ? src[static_cast<long long>(sy) * sstride + (sx + 1) * 3 + c] : 0;
        const int p2 = (x0ok && y1ok) ? src[static_cast<long long>(sy + 1) * sstride + sx * 3 + c] : 0;
        const int p3 = (x1ok && y1ok) ? src[static_cast<long long>(sy + 1) * sstride + (sx + 1) * 3 + c] : 0;
        v[c] = ((p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3 + (1 << 14)) >> 15) & 0xFF;
      }
    }
    if (swap_rb) { const int t = v[0]; v[0] = v[2]; v[2] = t; }
    if (out_u8) {
      uint8_t* o = out_u8 + ((static_cast<size_t>(face) * PRE_OUT + y) * PRE_OUT + x) * 3;
      o[0] = static_cast<uint8_t>(v[0]); o[1] = static_cast<uint8_t>(v[1]); o[2] = static_cast<uint8_t>(v[2]);
    }
    if (out_f16) store_s2d(static_cast<float>(v[0]), static_cast<float>(v[1]), static_cast<float>(v[2]), face, y, x, out_f16);
  }
}

// BilinearTab_i of OpenCV's remap, built exactly like initInterTab2D does (see oracle/warp_oracle.c for the fix-up quirk)
static void build_warp_tab(short* flat /* [1024 * 4 + 8] */) {
  for (int i = 0; i < 1024 * 4 + 8; ++i) flat[i] = 0;
  for (int i = 0; i < 32; ++i) {
    const float xi = static_cast<float>(i) * (1.f / 32);
    const float ty[2] = {1.f - xi, xi};
    for (int j = 0; j < 32; ++j) {
      const float xj = static_cast<float>(j) * (1.f / 32);
      const float tx[2] = {1.f - xj, xj};
      short* it = flat + (i * 32 + j) * 4;
      int isum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const int q = static_cast<int>(lrintf(ty[k1] * tx[k2] * 32768.f));
          it[k1 * 2 + k2] = static_cast<short>(q > 32767 ? 32767 : (q < -32768 ? -32768 : q));
          isum += it[k1 * 2 + k2];
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int Mk = 3, mk = 3;
        for (int idx = 3; idx <= 6; ++idx) {
          if (it[idx] < it[mk]) mk = idx;
          else if (it[idx] > it[Mk]) Mk = idx;
        }
        if (diff < 0) it[Mk] = static_cast<short>(it[Mk] - diff);
        else it[mk] = static_cast<short>(it[mk] - diff);
      }
    }
  }
}

}  // namespace fire

using namespace fire;

extern "C" int fire_align_warp(const uint8_t* frames, const int64_t* frame_desc, int n_frames, const double* matrices,
                               const int32_t* face_frame, int n_faces, int swap_rb, uint8_t* out_u8, void* out_f16,
                               fire_stream_t stream) {
  if (!frames || !frame_desc || !matrices || !face_frame) return fail(FIRE_ERR_ARG, "fire_align_warp: NULL argument");
  if (!out_u8 && !out_f16) return fail(FIRE_ERR_ARG, "fire_align_warp: no output requested");
  if (n_faces <= 0 || n_frames <= 0) return fail(FIRE_ERR_ARG, "fire_align_warp: n_faces=%d n_frames=%d", n_faces, n_frames);
  static bool tab_ready = false;
  if (!tab_ready) {
    short flat[1024 * 4 + 8];
    build_warp_tab(flat);
    FIRE_CUDA(cudaMemcpyToSymbol(g_warp_tab, flat, sizeof(short) * 1024 * 4));
    tab_ready = true;
  }
  dim3 grid(n_faces, PRE_OUT / PRE_ROWS_PER_BLOCK);
  align_warp_kernel<<<grid, PRE_OUT, 0, static_cast<cudaStream_t>(stream)>>>(frames, frame_desc, matrices, face_frame, swap_rb ? 1 : 0, out_u8,
                                                                              static_cast<__half*>(out_f16));
  FIRE_LAUNCH_CHECK("align_warp_kernel");
  count_launch();
  return FIRE_OK;
}

extern "C" int fire_preprocess(const uint8_t* frames, const int64_t* frame_desc, int n_frames, const int32_t* boxes_xywh,
                               const int32_t* box_frame, int n_boxes, int mode, void* out_f16, float* out_f32,
                               int32_t* box_status, fire_stream_t stream) {
  if (!frames || !frame_desc || !boxes_xywh || !box_frame) return fail(FIRE_ERR_ARG, "fire_preprocess: NULL argument");
  if (!out_f16 && !out_f32) return fail(FIRE_ERR_ARG, "fire_preprocess: no output requested");
  if (n_boxes <= 0 || n_frames <= 0) return fail(FIRE_ERR_ARG, "fire_preprocess: n_boxes=%d n_frames=%d", n_boxes, n_frames);
  const int swap = (mode & FIRE_PRE_FLAG_SWAP_RB) ? 1 : 0;
  const int m = mode & 15;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m == FIRE_PRE_REFERENCE) {
    dim3 grid(n_boxes, PRE_OUT / PRE_ROWS_PER_BLOCK);
    static bool attr_done[FIRE_MAX_DEVICES] = {};
    int dev = 0;
    FIRE_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < FIRE_MAX_DEVICES && !attr_done[dev]) {
      FIRE_CUDA(cudaFuncSetAttribute(preprocess_reference_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PRE_STAGE_BYTES));
      attr_done[dev] = true;
    }
    preprocess_reference_kernel<<<grid, PRE_THREADS, PRE_STAGE_BYTES, st>>>(frames, frame_desc, boxes_xywh, box_frame, swap,
                                                              static_cast<__half*>(out_f16), out_f32, box_status);
  } else if (m == FIRE_PRE_NORTHSTAR) {
    preprocess_northstar_kernel<<<n_boxes, NS_THREADS, 0, st>>>(frames, frame_desc, boxes_xywh, box_frame, swap,
                                                                static_cast<__half*>(out_f16), out_f32, box_status);
  } else {
    return fail(FIRE_ERR_ARG, "fire_preprocess: unknown mode %d", mode);
  }
  FIRE_LAUNCH_CHECK("preprocess kernel");
  count_launch();
  return FIRE_OK;
}
