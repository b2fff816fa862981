// pool_conv_fused.cuh - MaxPool_3a (3x3 / stride 2, valid) + Conv2d_3b (1x1, 64 -> 80, BN folded, ReLU) as ONE tcgen05 launch.
//
// Reference graph (stem; executed by onnxruntime at facenet_gpu.py:127):
//     p = maxpool3x3s2(x)   x: 77 x 77 x 64 (the output of Conv2d_2b, the largest tensor of the network: 194 MB at B = 256)
//     y = relu(bn(conv1x1 64 -> 80 (p)))                                                   y: 38 x 38 x 80
// As two launches the pooled tensor (47 MB) is written and read back, and both launches are HBM-bound (55 + 27 us).  Here the
// pooled tile never leaves the SM.  A first attempt (round 1, profiles/r01_probe_notes.md) let the conv's 256 gather threads
// load the nine window taps from global memory: too few bytes in flight to reach the HBM rate.  This kernel lets the TMA do
// the fetching: one 4-D box (64 ch x 77 px x 7 rows) = the input of three pooled rows lands in shared memory per tile, eight
// warps reduce the 3 x 3 windows FROM SHARED MEMORY into the swizzled A operand [114 positions x 64 ch], one elected thread
// issues the four K = 16 MMAs (N = 80), four warps convert and stage, TMA stores write the 114 x 80 tile back.
//
// Tile = (image, block of 3 pooled rows): 13 blocks per image (the last one has 2 rows; the TMA zero-fills the rows past the
// image, the store map clips the positions past it).  Persistent CTAs; patch stages, A operand and accumulator are all
// double-buffered, so the pooling of tile t + 1 overlaps the MMA / epilogue / store of tile t and the TMA load of tile t + 2.
#pragma once

#include "block17_fused.cuh"

namespace fire {

constexpr int PC_CIN = 64, PC_COUT = 80;
constexpr int PC_IN = 77, PC_OUT = 38;                 // input / pooled rows and columns
constexpr int PC_R = 3;                                // pooled rows per tile
constexpr int PC_POS = PC_R * PC_OUT;                  // 114 positions = rows of the UMMA M tile that carry data
constexpr int PC_IN_ROWS = 2 * PC_R + 1;               // 7 input rows per tile
constexpr int PC_ROW_BLOCKS = (PC_OUT + PC_R - 1) / PC_R;   // 13
constexpr uint32_t PC_PATCH_BYTES = PC_IN_ROWS * PC_IN * 128;          // 68992: one 128-byte swizzled row per input pixel
constexpr uint32_t PC_PATCH_STRIDE = (PC_PATCH_BYTES + 1023) & ~1023u;  // 69632
constexpr int PC_POOL_WARPS = 8, PC_EPI_WARPS = 4;
constexpr int PC_THREADS = 32 * (2 + PC_EPI_WARPS + PC_POOL_WARPS);     // warp 0 producer, 1 MMA, 2-5 epilogue, 6-13 pooling

constexpr uint32_t PC_PATCH = 0;
constexpr uint32_t PC_A = PC_PATCH + 2 * PC_PATCH_STRIDE;               // 2 x [128 x 64] SWIZZLE_128B
constexpr uint32_t PC_W = PC_A + 2 * 16384;                             // [80 x 64] SWIZZLE_128B
constexpr uint32_t PC_OUTA = PC_W + 10240;                              // staging, channels 0-63: [128 rows x 128 B] SWIZZLE_128B
constexpr uint32_t PC_OUTB = PC_OUTA + 16384;                           // staging, channels 64-79: [128 rows x 32 B] SWIZZLE_32B
constexpr uint32_t PC_BIAS = PC_OUTB + 4096;                            // [80] fp32
constexpr uint32_t PC_BARS = PC_BIAS + 512;
constexpr uint32_t PC_SMEM = PC_BARS + 256 + 1024;
static_assert(PC_A % 1024 == 0 && PC_W % 1024 == 0 && PC_OUTA % 1024 == 0 && PC_OUTB % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(PC_SMEM <= 232448, "shared memory budget");

struct PoolConvParams {
  CUtensorMap in_map;       // x [B][77][pitch][64]: box (64 ch, 77 px, 7 rows, 1 image), SWIZZLE_128B
  CUtensorMap w_map;        // weights [80][64] K-major: box 64 x 80
  CUtensorMap out_a;        // y as {80 ch, 1444 positions, B}: box (64 ch, 114 positions, 1) - channels 0-63
  CUtensorMap out_b;        // same tensor: box (16 ch, 114 positions, 1) - channels 64-79
  const float* bias;        // [80]
  int n_images, pdl;
};

__device__ __forceinline__ uint4 hmax2x4(uint4 a, uint4 b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__global__ void __launch_bounds__(PC_THREADS, 1)
pool_conv_fused_kernel(const __grid_constant__ PoolConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PC_BARS);
  uint64_t* p_full = bars;              // [2] patch stage landed (TMA)
  uint64_t* p_empty = p_full + 2;       // [2] patch stage read by the eight pooling warps
  uint64_t* a_full = p_empty + 2;       // [2] A operand written by the eight pooling warps
  uint64_t* a_empty = a_full + 2;       // [2] ... read by the MMAs
  uint64_t* acc_full = a_empty + 2;     // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2] accumulator read by the four epilogue warps
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_images * PC_ROW_BLOCKS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.in_map); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.out_a); tma_prefetch_desc(&p.out_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], PC_POOL_WARPS);
        mbar_init(&a_full[s], PC_POOL_WARPS); mbar_init(&a_empty[s], 1);
        mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PC_EPI_WARPS);
      }
      mbar_init(w_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, 256u);
  }
  if (threadIdx.x < PC_COUT) reinterpret_cast<float*>(smem + PC_BIAS)[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) pdl_launch_dependents();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: weights once, then one patch per tile
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, PC_COUT * 128);
      tma_load_2d_hint(smem + PC_W, &p.w_map, w_full, 0, 0, kEvictLast);
    }
    __syncwarp();
    if (p.pdl) pdl_wait();                                        // x is the previous layer's output
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int img = tile / PC_ROW_BLOCKS, rb = tile - img * PC_ROW_BLOCKS;
      const int s = k & 1;
      mbar_wait(&p_empty[s], ((k >> 1) & 1) ^ 1, 90);
      if (elect_one()) {
        mbar_arrive_expect_tx(&p_full[s], PC_PATCH_BYTES);
        tma_load_4d(smem + PC_PATCH + s * PC_PATCH_STRIDE, &p.in_map, &p_full[s], 0, 0, 2 * PC_R * rb, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: D[128 x 80] = A[128 x 64] * W^T
    const uint32_t idesc = umma_idesc_f16(CONV_BM, PC_COUT);
    mbar_wait(w_full, 0, 91);
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t par = (k >> 1) & 1;
      mbar_wait(&acc_empty[s], par ^ 1, 92);
      mbar_wait(&a_full[s], par, 93);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = sbase + PC_A + static_cast<uint32_t>(s) * 16384u, b0 = sbase + PC_W;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + static_cast<uint32_t>(s * 128), umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc, kk != 0 ? 1u : 0u);
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[s]);
      }
      __syncwarp();
    }
  } else if (warp < 2 + PC_EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue: + bias, ReLU, fp16 -> staging -> TMA stores
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_bias = sbase + PC_BIAS;
    const uint32_t row_a = sbase + PC_OUTA + static_cast<uint32_t>(m * 128), swz_a = m & 7;
    const uint32_t row_b = sbase + PC_OUTB + static_cast<uint32_t>(m * 32), swz_b = (m >> 2) & 1;
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int img = tile / PC_ROW_BLOCKS, rb = tile - img * PC_ROW_BLOCKS;
      const int s = k & 1;
      mbar_wait(&acc_full[s], (k >> 1) & 1, 94);
      tc_fence_after();
      uint32_t buf[2][16];
      __syncwarp();
      tmem_ld_32x16(tq + static_cast<uint32_t>(s * 128), buf[0]);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t (&r)[16] = buf[c & 1];
        tmem_ld_wait(r);
        if (c + 1 < 5) tmem_ld_32x16(tq + static_cast<uint32_t>(s * 128 + 16 * (c + 1)), buf[(c + 1) & 1]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b = as_f4(lds128(s_bias + static_cast<uint32_t>((16 * c + 4 * e) * 4)));
          r[4 * e] = __float_as_uint(__uint_as_float(r[4 * e]) + b.x);
          r[4 * e + 1] = __float_as_uint(__uint_as_float(r[4 * e + 1]) + b.y);
          r[4 * e + 2] = __float_as_uint(__uint_as_float(r[4 * e + 2]) + b.z);
          r[4 * e + 3] = __float_as_uint(__uint_as_float(r[4 * e + 3]) + b.w);
        }
        if (c < 4) conv_stage_chunk(r, true, row_a, static_cast<uint32_t>(2 * c), swz_a);
        else conv_stage_chunk(r, true, row_b, 0u, swz_b);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
      fence_proxy_async_smem();
      named_bar_sync(1, PC_EPI_WARPS * 32);                      // the whole tile is staged
      if (warp == 2) {
        if (elect_one()) {
          tma_store_3d(&p.out_a, sbase + PC_OUTA, 0, rb * PC_POS, img);
          tma_store_3d(&p.out_b, sbase + PC_OUTB, 64, rb * PC_POS, img);
          bulk_commit_group();
          bulk_wait_read_all();                                   // staging may be overwritten
        }
        __syncwarp();
      }
      named_bar_sync(1, PC_EPI_WARPS * 32);
    }
    if (warp == 2) {
      if (elect_one()) bulk_wait_all();                           // stores complete before the CTA exits
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- pooling: 3 x 3 / 2 windows of the patch -> A operand
    const int t = threadIdx.x - 32 * (2 + PC_EPI_WARPS);          // 0..255
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t par = (k >> 1) & 1;
      mbar_wait(&p_full[s], par, 95);
      mbar_wait(&a_empty[s], par ^ 1, 96);
      const uint32_t patch = sbase + PC_PATCH + static_cast<uint32_t>(s) * PC_PATCH_STRIDE;
      const uint32_t a_tile = sbase + PC_A + static_cast<uint32_t>(s) * 16384u;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int i = t + it * 256;                               // (position, 16-byte channel chunk)
        if (i < PC_POS * 8) {
          const int pos = i >> 3, c = i & 7;
          const int pr = pos / PC_OUT, px = pos - pr * PC_OUT;
          const int q0 = 2 * pr * PC_IN + 2 * px;                // first pixel of the window inside the patch
          uint4 v = lds128(patch + static_cast<uint32_t>(q0 * 128 + ((c ^ (q0 & 7)) << 4)));
#pragma unroll
          for (int w = 1; w < 9; ++w) {
            const int q = q0 + (w / 3) * PC_IN + (w % 3);
            v = hmax2x4(v, lds128(patch + static_cast<uint32_t>(q * 128 + ((c ^ (q & 7)) << 4))));
          }
          sts128(a_tile + static_cast<uint32_t>(pos * 128 + ((c ^ (pos & 7)) << 4)), v);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a_full[s]); mbar_arrive(&p_empty[s]); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, 256u);
  }
}

}  // namespace fire
