// block17_fused.cuh - the ten Block17 (Inception-ResNet-B) blocks of FaceNet as ONE persistent tcgen05 kernel.
//
// Reference graph (executed by onnxruntime at facenet_gpu.py:127), per block, on an 8 x 8 x 896 map:
//     b0  = relu(bn(conv1x1 896->128 (x)))            b1a = relu(bn(conv1x1 896->128 (x)))
//     b1b = relu(bn(conv1x7 128->128 (b1a)))          b1c = relu(bn(conv7x1 128->128 (b1b)))
//     y   = relu(x + 0.1 * (conv1x1 256->896 ([b0 | b1c]) + bias))
// Layer by layer (conv_igemm_kernel x 4 per block) the 8 x 8 stage is ONE wave of 128 tiles, so every layer pays its
// full pipeline fill/drain (~5 us of a 12 us layer), re-reads the 128 x 896 tile and round-trips the intermediates
// through L2.  None of these convolutions mixes images, so a CTA that owns TWO images (128 positions = one UMMA M
// tile) can run the whole ten-block chain on its own: no grid-wide dependency is left, only CTA-local ones.
//
// Per CTA and block:
//   H    D[128 x 256] = x[128 x 896] * Wh^T          x K-blocks by TMA (6-slot ring X), weights from the stream
//        epilogue: +bias, ReLU, fp16 ->  b1a into R1 (x-major rows, see below),  b0 into R0 (natural rows)
//   1x7  D[128 x 128] = sum_s R1[rows + 16(s-3)] * W_s^T     A operand = a row-shifted window of R1, no im2col
//        epilogue -> b1b into R2 (y-major rows)
//   7x1  D[128 x 128] = sum_r R2[rows + 16(r-3)] * W_r^T
//        epilogue -> b1c into R3 (= the data rows of R1, natural order)
//   up   D[128 x 896] = [R0 | R3] * Wu^T in N tiles 256, 256, 256, 128 (two TMEM buffers)
//        epilogue: + bias + x (residual read straight from L2), ReLU, fp16 -> staging -> TMA store into y
//   the next block's H reads y through the TMA as soon as this CTA's stores have completed (y_done).
//
// Row orders.  R1 keeps b1a as row = x*16 + img*8 + y with 48 zero rows before and after: the input of tap s of the
// 1 x 7 conv for ALL 128 outputs is the same buffer shifted by 16(s-3) rows, and shifts past the image edge land in the
// zero rows - every MMA is a full, fully useful M = 128.  R2 is the same with the roles of x and y exchanged (7 x 1).
// The epilogues do the re-ordering for free: each thread owns one accumulator row and writes it wherever it belongs.
//
// Weights are repacked once on the host (facenet_engine.cu: pack_block17_stream) into the exact order the MMA warp
// consumes them, as 16 KB units that are byte images of the swizzled shared-memory operand ([256 x 32] SWIZZLE_64B
// for the N = 256 GEMMs, [128 x 64] SWIZZLE_128B for the N = 128 ones): the weight producer is a linear stream of
// cp.async.bulk copies through a 6-slot ring, with no tensor map and no per-layer prologue.  84 units per block.
//
// Shared memory (227 KB): [pad][R1 k0][pad][R1 k1][pad][R2 k0][pad][R2 k1][pad] (pads 48 rows x 128 B, zero),
// R0 (32 KB), ring U (6 x 16 KB), barriers, `up` bias.  While H runs, R0/R1/R2 hold nothing live: their 16 KB data
// regions ARE ring X.  During `up`, R2's data regions are the epilogue warps' transposition scratch.
// TMEM (512 columns): H -> [0,256), 1x7 -> [256,384), 7x1 -> [384,512), up tiles alternate [0,256) / [256,512).
//
// Warp roles: 0 and 10 weight stream (unit u belongs to issuer u % 2: the wait -> expect_tx -> copy chain of one
// thread costs ~700 cycles per unit, issuers in different warps overlap), 1 MMA issuer, 2-9 epilogue (two per TMEM
// lane quarter), 11 activation loads.
#pragma once

#include "conv_igemm.cuh"

namespace fire {

constexpr int B17_C = 896;                 // channels of x / y
constexpr int B17_MID = 128;
constexpr int B17_UNIT = 16384;
constexpr int B17_U_SLOTS = 6;
constexpr int B17_X_SLOTS = 6;
constexpr int B17_XK = B17_C / 64;         // 14 activation K-blocks per block
constexpr int B17_PAD = 48 * 128;          // 6144 B of zero rows
constexpr int B17_UNITS_PER_BLOCK = 28 + 14 + 14 + 28;     // 84
constexpr int B17_BIAS_PER_BLOCK = 256 + 128 + 128 + 896;  // 1408 floats: [heads | 1x7 | 7x1 | up]
constexpr int B17_MAX_BLOCKS = 10;
#ifdef FIRE_B200_SKIP_EXPERIMENTS
constexpr int B17_DBG_MASK = 3;
#else
constexpr int B17_DBG_MASK = 0;            // the shipped library cannot skip residual loads or output stores
#endif
constexpr int B17_TRACE_SLOTS = 24;         // per CTA and block: 0-7 MMA warp, 8-23 epilogue warp 2
constexpr int B17_THREADS = 32 * 12;        // 12 warps: the register file gives each thread 168 registers (14 warps: 128, with spills)
constexpr int B17_W_ISSUERS = 2;           // warps 0, 10: unit u is issued by issuer u % 2 (2 divides the ring: a slot has one owner)
constexpr int B17_X_ISSUERS = 1;           // warp 11

// shared-memory offsets (from the 1024-aligned base)
constexpr uint32_t B17_R1K0 = B17_PAD;                         //  6144
constexpr uint32_t B17_R1K1 = B17_R1K0 + B17_UNIT + B17_PAD;   // 28672
constexpr uint32_t B17_R2K0 = B17_R1K1 + B17_UNIT + B17_PAD;   // 51200
constexpr uint32_t B17_R2K1 = B17_R2K0 + B17_UNIT + B17_PAD;   // 73728
constexpr uint32_t B17_R0 = B17_R2K1 + B17_UNIT + B17_PAD;     // 96256
constexpr uint32_t B17_U = B17_R0 + 2 * B17_UNIT;              // 129024
constexpr uint32_t B17_BARS = B17_U + B17_U_SLOTS * B17_UNIT;  // 227328
constexpr uint32_t B17_BIAS_UP = B17_BARS + 512;               // [896] fp32 bias of the current block's `up` conv
constexpr uint32_t B17_SMEM = B17_BIAS_UP + 896 * 4 + 1024;    // + alignment slack = 232448
static_assert(B17_R1K0 % 1024 == 0 && B17_R1K1 % 1024 == 0 && B17_R2K0 % 1024 == 0 && B17_R2K1 % 1024 == 0 && B17_R0 % 1024 == 0 &&
              B17_U % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(B17_SMEM <= 232448, "shared memory budget");

struct B17Params {
  CUtensorMap xmap[B17_MAX_BLOCKS + 1];   // x_0 .. x_n: [M][896] fp16, box 64 columns x 128 rows, SWIZZLE_128B (loads and stores)
  const __half* xptr[B17_MAX_BLOCKS + 1];
  const uint8_t* wstream;                 // n_blocks x 84 units
  const float* bias;                      // n_blocks x 1408
  int n_blocks, M_total, n_tiles, pdl;
  int dbg;                                // FIRE_B200_SKIP_EXPERIMENTS builds only (B17_DBG_MASK is 0 otherwise): 1 = no residual loads, 2 = no y stores
  long long* trace;                       // optional: [gridDim.x][B17_MAX_BLOCKS][B17_TRACE_SLOTS] globaltimer stamps
};

__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictLast)
               : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {      // L2-only load: the data was written by this CTA's TMA stores
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 16 accumulator columns (+ bias) -> ReLU -> fp16 -> two 16-byte units of a 128-byte-swizzled row
__device__ __forceinline__ void b17_store_chunk(uint32_t (&r)[16], float4 b0, float4 b1, float4 b2, float4 b3, uint32_t row_addr, uint32_t u0, uint32_t swz) {
  const float4 bb[4] = {b0, b1, b2, b3};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 b = bb[e];
    r[4 * e] = __float_as_uint(__uint_as_float(r[4 * e]) + b.x);
    r[4 * e + 1] = __float_as_uint(__uint_as_float(r[4 * e + 1]) + b.y);
    r[4 * e + 2] = __float_as_uint(__uint_as_float(r[4 * e + 2]) + b.z);
    r[4 * e + 3] = __float_as_uint(__uint_as_float(r[4 * e + 3]) + b.w);
  }
  conv_stage_chunk(r, true, row_addr, u0, swz);
}
// 64 accumulator columns of this thread's row -> + bias, ReLU, fp16 -> one swizzled 128-byte smem row; the TMEM load of
// chunk c + 1 is in flight while chunk c is converted
__device__ __forceinline__ void b17_epi_row64(uint32_t taddr, const float4 (&bq)[16], uint32_t row_addr, uint32_t swz) {
  uint32_t a[16], b[16];
  __syncwarp();
  tmem_ld_32x16(taddr, a);
  tmem_ld_wait(a);
  tmem_ld_32x16(taddr + 16, b);
  b17_store_chunk(a, bq[0], bq[1], bq[2], bq[3], row_addr, 0u, swz);
  tmem_ld_wait(b);
  tmem_ld_32x16(taddr + 32, a);
  b17_store_chunk(b, bq[4], bq[5], bq[6], bq[7], row_addr, 2u, swz);
  tmem_ld_wait(a);
  tmem_ld_32x16(taddr + 48, b);
  b17_store_chunk(a, bq[8], bq[9], bq[10], bq[11], row_addr, 4u, swz);
  tmem_ld_wait(b);
  b17_store_chunk(b, bq[12], bq[13], bq[14], bq[15], row_addr, 6u, swz);
}
__device__ __forceinline__ void b17_pack_chunk(uint32_t (&r)[16], float4 b0, float4 b1, float4 b2, float4 b3, uint4& lo, uint4& hi);
// Same for the two TRANSPOSING epilogues (H -> R1, 1x7 -> R2): the eight lanes of a quarter warp own rows that are 16
// rows apart, i.e. the same swizzle phase, so writing "unit u of my row" from all lanes is an 8-way bank conflict
// (measured: 1.4 us instead of 0.4 us per phase).  Here lane l writes unit (i + l) & 7 at step i: the row is packed into
// registers first and rotated by l & 7 with a three-stage select network.
__device__ __forceinline__ void b17_sel_rot(uint4 (&v)[8], bool on, int by) {
  uint4 t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a = v[i], b = v[(i + by) & 7];
    t[i] = make_uint4(on ? b.x : a.x, on ? b.y : a.y, on ? b.z : a.z, on ? b.w : a.w);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = t[i];
}
__device__ __forceinline__ void b17_epi_row64_rot(uint32_t taddr, const float4 (&bq)[16], uint32_t row_addr, uint32_t swz, int lane) {
  uint32_t a[16], b[16];
  uint4 pk[8];
  __syncwarp();
  tmem_ld_32x16(taddr, a);
  tmem_ld_wait(a);
  tmem_ld_32x16(taddr + 16, b);
  b17_pack_chunk(a, bq[0], bq[1], bq[2], bq[3], pk[0], pk[1]);
  tmem_ld_wait(b);
  tmem_ld_32x16(taddr + 32, a);
  b17_pack_chunk(b, bq[4], bq[5], bq[6], bq[7], pk[2], pk[3]);
  tmem_ld_wait(a);
  tmem_ld_32x16(taddr + 48, b);
  b17_pack_chunk(a, bq[8], bq[9], bq[10], bq[11], pk[4], pk[5]);
  tmem_ld_wait(b);
  b17_pack_chunk(b, bq[12], bq[13], bq[14], bq[15], pk[6], pk[7]);
  const int s = lane & 7;
  b17_sel_rot(pk, (s & 1) != 0, 1);
  b17_sel_rot(pk, (s & 2) != 0, 2);
  b17_sel_rot(pk, (s & 4) != 0, 4);                  // pk[i] = unit (i + s) & 7
#pragma unroll
  for (int i = 0; i < 8; ++i) sts128(row_addr + (((static_cast<uint32_t>(i + s) & 7u) ^ swz) << 4), pk[i]);
}
// same, packed into registers (the `up` epilogue stores to global memory)
__device__ __forceinline__ void b17_pack_chunk(uint32_t (&r)[16], float4 b0, float4 b1, float4 b2, float4 b3, uint4& lo, uint4& hi) {
  const float4 bb[4] = {b0, b1, b2, b3};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 b = bb[e];
    r[4 * e] = __float_as_uint(__uint_as_float(r[4 * e]) + b.x);
    r[4 * e + 1] = __float_as_uint(__uint_as_float(r[4 * e + 1]) + b.y);
    r[4 * e + 2] = __float_as_uint(__uint_as_float(r[4 * e + 2]) + b.z);
    r[4 * e + 3] = __float_as_uint(__uint_as_float(r[4 * e + 3]) + b.w);
  }
  lo = make_uint4(cvt_pack_relu(r[0], r[1]), cvt_pack_relu(r[2], r[3]), cvt_pack_relu(r[4], r[5]), cvt_pack_relu(r[6], r[7]));
  hi = make_uint4(cvt_pack_relu(r[8], r[9]), cvt_pack_relu(r[10], r[11]), cvt_pack_relu(r[12], r[13]), cvt_pack_relu(r[14], r[15]));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 as_f4(uint4 v) {
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ void st_global_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void b17_add_res(uint32_t (&r)[16], uint4 lo, uint4 hi) {
  auto add2 = [&](int e, uint32_t w) {
    __half2 hh;
    memcpy(&hh, &w, 4);
    const float2 f = __half22float2(hh);
    r[2 * e] = __float_as_uint(__uint_as_float(r[2 * e]) + f.x);
    r[2 * e + 1] = __float_as_uint(__uint_as_float(r[2 * e + 1]) + f.y);
  };
  add2(0, lo.x); add2(1, lo.y); add2(2, lo.z); add2(3, lo.w);
  add2(4, hi.x); add2(5, hi.y); add2(6, hi.z); add2(7, hi.w);
}

#define B17_TRACE(blk_, slot_) do { if (p.trace && lane == 0) p.trace[(static_cast<size_t>(blockIdx.x) * B17_MAX_BLOCKS + (blk_)) * B17_TRACE_SLOTS + (slot_)] = globaltimer_ns(); } while (0)

__global__ void __launch_bounds__(B17_THREADS, 1)
block17_fused_kernel(const __grid_constant__ B17Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B17_BARS);
  uint64_t* u_full = bars;                       // [6]
  uint64_t* u_empty = u_full + B17_U_SLOTS;      // [6]
  uint64_t* x_full = u_empty + B17_U_SLOTS;      // [6]
  uint64_t* x_empty = x_full + B17_X_SLOTS;      // [6]
  uint64_t* accH_full = x_empty + B17_X_SLOTS;
  uint64_t* acc17_full = accH_full + 1;
  uint64_t* acc71_full = acc17_full + 1;
  uint64_t* accU_full = acc71_full + 1;          // [2]
  uint64_t* accU_empty = accU_full + 2;          // [2]
  uint64_t* r1_ready = accU_empty + 2;
  uint64_t* r0_ready = r1_ready + 1;
  uint64_t* r2_ready = r0_ready + 1;
  uint64_t* r3_ready = r2_ready + 1;
  uint64_t* y_done = r3_ready + 1;               // the epilogue warps have written this block's y (= the next block's x)
  uint64_t* y_part = y_done + 1;                 // [3] ... N tile t of it (columns 256 t .. 256 t + 255): the next block's x K-blocks 4 t .. 4 t + 3
  uint64_t* up_done = y_part + 3;               // every MMA of this block has completed (R0 / R1 are free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(up_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j <= p.n_blocks; ++j) tma_prefetch_desc(&p.xmap[j]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < B17_U_SLOTS; ++s) { mbar_init(&u_full[s], 1); mbar_init(&u_empty[s], 1); }
      for (int s = 0; s < B17_X_SLOTS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
      mbar_init(accH_full, 1); mbar_init(acc17_full, 1); mbar_init(acc71_full, 1);
      for (int b = 0; b < 2; ++b) { mbar_init(&accU_full[b], 1); mbar_init(&accU_empty[b], CONV_EPI_WARPS); }
      mbar_init(r1_ready, CONV_EPI_WARPS); mbar_init(r0_ready, CONV_EPI_WARPS);
      mbar_init(r2_ready, CONV_EPI_WARPS); mbar_init(r3_ready, CONV_EPI_WARPS);
      mbar_init(y_done, CONV_EPI_WARPS);
      for (int t = 0; t < 3; ++t) mbar_init(&y_part[t], CONV_EPI_WARPS);
      mbar_init(up_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, 512u);
  }
  {
    // the five zero pads (nobody ever writes them again)
    for (int i = threadIdx.x; i < 5 * (B17_PAD / 16); i += B17_THREADS) {
      const int q = i / (B17_PAD / 16), o = i - q * (B17_PAD / 16);
      reinterpret_cast<uint4*>(smem + q * (B17_UNIT + B17_PAD))[o] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) pdl_launch_dependents();

  const int my_tiles = (p.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // R1 / R2 K-block regions are B17_UNIT + B17_PAD apart
  auto r1k = [sbase](int kb) { return sbase + B17_R1K0 + static_cast<uint32_t>(kb) * (B17_UNIT + B17_PAD); };
  auto r2k = [sbase](int kb) { return sbase + B17_R2K0 + static_cast<uint32_t>(kb) * (B17_UNIT + B17_PAD); };
  const uint32_t r0a = sbase + B17_R0;
  const uint32_t ubase = sbase + B17_U;

  const int w_issuer = warp == 0 ? 0 : warp == 10 ? 1 : -1;
  const int x_issuer = warp == 11 ? 0 : -1;
  if (w_issuer >= 0) {
    // ---------------------------------------------------------------- weight stream (independent of the previous layer)
    const int per_tile = p.n_blocks * B17_UNITS_PER_BLOCK;
    const int total = my_tiles * per_tile;
    int slot = w_issuer, src = w_issuer;
    uint32_t ph = 0;
    for (int u = w_issuer; u < total; u += B17_W_ISSUERS) {
      mbar_wait(&u_empty[slot], ph ^ 1, 31);
      if (elect_one()) {
        mbar_arrive_expect_tx(&u_full[slot], B17_UNIT);
        bulk_copy_g2s(ubase + static_cast<uint32_t>(slot) * B17_UNIT, p.wstream + static_cast<size_t>(src) * B17_UNIT, B17_UNIT, &u_full[slot]);
      }
      __syncwarp();
      slot += B17_W_ISSUERS;
      if (slot >= B17_U_SLOTS) { slot -= B17_U_SLOTS; ph ^= 1; }
      src += B17_W_ISSUERS;
      if (src >= per_tile) src -= per_tile;
    }
  } else if (x_issuer >= 0) {
    // ---------------------------------------------------------------- activation K-blocks of H (ring X = the idle R0/R1/R2 data regions)
    if (p.pdl) pdl_wait();
    int blk = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int m0 = tile * CONV_BM;
      for (int j = 0; j < p.n_blocks; ++j, ++blk) {
        // Ring X: K-blocks 0-11 cycle through R0 / R1's data rows (4 slots), which are free as soon as the previous block's
        // last `up` MMA has completed; K-block i only needs the N tile i / 4 of the previous block's y, which its epilogue
        // publishes tile by tile (y_part): H starts on the first tiles while that epilogue is still busy with the last ones.
        // K-blocks 12, 13 go to R2's data rows (the epilogue's scratch) once all of y is written (y_done).
        if (blk > 0) mbar_wait(up_done, (blk - 1) & 1, 32);
        for (int i = x_issuer; i < B17_XK; i += B17_X_ISSUERS) {
          const int k = i >> 2;                                 // use of the slot within this block (0..2), or 3 for the R2 slots
          const int xs = i < 12 ? (i & 3) : i - 8;              // 0..3, 4, 5
          if (blk > 0 && (i & 3) == 0) {
            if (i < 12) mbar_wait(&y_part[k], (blk - 1) & 1, 32); else mbar_wait(y_done, (blk - 1) & 1, 32);
            fence_proxy_async_all();
          }
          const uint32_t xoff = xs == 0 ? B17_R0 : xs == 1 ? B17_R0 + B17_UNIT : xs == 2 ? B17_R1K0 : xs == 3 ? B17_R1K1 : xs == 4 ? B17_R2K0 : B17_R2K1;
          // slots 0-3 are filled three times per block (fill number 3 blk + k), slots 4 and 5 once
          if (i >= 4 && i < 12) mbar_wait(&x_empty[xs], (static_cast<uint32_t>(blk) + k - 1) & 1, 33);
          if (elect_one()) {
            mbar_arrive_expect_tx(&x_full[xs], B17_UNIT);
            tma_load_2d(smem + xoff, &p.xmap[j], &x_full[xs], i * 64, m0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc256 = umma_idesc_f16(CONV_BM, 256), idesc128 = umma_idesc_f16(CONV_BM, 128);
    int us = 0;
    uint32_t uph = 0;
    int blk = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_blocks; ++j, ++blk) {
        const uint32_t bpar = blk & 1;
        // ---- H: D[0,256) = x * Wh^T
        if (blk > 0) mbar_wait(&accU_empty[0], 1, 34);        // `up` tile 2 of the previous block has left TMEM [0,256)
        tc_fence_after();
        B17_TRACE(j, 0);
        for (int i = 0; i < B17_XK; ++i) {
          const int k = i >> 2;
          const int xs = i < 12 ? (i & 3) : i - 8;
          const uint32_t xaddr = xs == 0 ? r0a : xs == 1 ? r0a + B17_UNIT : xs == 2 ? r1k(0) : xs == 3 ? r1k(1) : xs == 4 ? r2k(0) : r2k(1);
          mbar_wait(&x_full[xs], i < 12 ? (bpar + k) & 1 : bpar, 35);
          if (i == 0) B17_TRACE(j, 1);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&u_full[us], uph, 36);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b0 = ubase + static_cast<uint32_t>(us) * B17_UNIT;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_f16(tmem_base, umma_desc_sw128(xaddr + (half * 2 + kk) * 32), umma_desc_swz(b0 + kk * 32, 64), idesc256, (i | half | kk) != 0 ? 1u : 0u);
              umma_commit(&u_empty[us]);
              if (half == 1) umma_commit(&x_empty[xs]);
            }
            __syncwarp();
            if (++us == B17_U_SLOTS) { us = 0; uph ^= 1; }
          }
        }
        if (elect_one()) umma_commit(accH_full);
        __syncwarp();
        B17_TRACE(j, 2);
        // ---- 1x7 and 7x1: D = sum over taps of a row-shifted window of R1 / R2
#pragma unroll 1
        for (int conv = 0; conv < 2; ++conv) {
          mbar_wait(conv == 0 ? r1_ready : r2_ready, bpar, 37);
          if (conv == 0 && blk > 0) mbar_wait(&accU_empty[1], 1, 38);   // `up` tile 3 (and 1) of the previous block has left TMEM [256,512)
          tc_fence_after();
          if (conv == 0) B17_TRACE(j, 3);
          const uint32_t d = tmem_base + (conv == 0 ? 256u : 384u);
          for (int s = 0; s < 7; ++s) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              mbar_wait(&u_full[us], uph, 39);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a0 = (conv == 0 ? r1k(kb) : r2k(kb)) + static_cast<uint32_t>((s - 3) * 2048);
                const uint32_t b0 = ubase + static_cast<uint32_t>(us) * B17_UNIT;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(d, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc128, (s | kb | kk) != 0 ? 1u : 0u);
                umma_commit(&u_empty[us]);
              }
              __syncwarp();
              if (++us == B17_U_SLOTS) { us = 0; uph ^= 1; }
            }
          }
          if (elect_one()) umma_commit(conv == 0 ? acc17_full : acc71_full);
          __syncwarp();
        }
        // ---- up: D = [R0 | R3] * Wu^T, N tiles 256, 256, 256, 128
        mbar_wait(r0_ready, bpar, 40);
        mbar_wait(r3_ready, bpar, 41);
        tc_fence_after();
        B17_TRACE(j, 4);
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
          if (t == 2) B17_TRACE(j, 6);
          if (t >= 2) { mbar_wait(&accU_empty[t & 1], 0, 42); tc_fence_after(); }
          if (t == 2) B17_TRACE(j, 7);
          const uint32_t d = tmem_base + static_cast<uint32_t>((t & 1) * 256);
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t a0 = kb < 2 ? r0a + static_cast<uint32_t>(kb) * B17_UNIT : r1k(kb - 2);
            if (t < 3) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                mbar_wait(&u_full[us], uph, 43);
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t b0 = ubase + static_cast<uint32_t>(us) * B17_UNIT;
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk)
                    umma_f16(d, umma_desc_sw128(a0 + (half * 2 + kk) * 32), umma_desc_swz(b0 + kk * 32, 64), idesc256, (kb | half | kk) != 0 ? 1u : 0u);
                  umma_commit(&u_empty[us]);
                }
                __syncwarp();
                if (++us == B17_U_SLOTS) { us = 0; uph ^= 1; }
              }
            } else {
              mbar_wait(&u_full[us], uph, 44);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b0 = ubase + static_cast<uint32_t>(us) * B17_UNIT;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(d, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc128, (kb | kk) != 0 ? 1u : 0u);
                umma_commit(&u_empty[us]);
              }
              __syncwarp();
              if (++us == B17_U_SLOTS) { us = 0; uph ^= 1; }
            }
          }
          if (elect_one()) {
            umma_commit(&accU_full[t & 1]);
            if (t == 3) umma_commit(up_done);
          }
          __syncwarp();
        }
        B17_TRACE(j, 5);
      }
    }
  } else if (warp >= CONV_FIRST_EPI_WARP && warp < CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue (8 warps: TMEM lane quarter x column half)
    const int quarter = warp & 3, h = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const int r = quarter * 32 + lane;                          // accumulator row of this thread
    const int et = threadIdx.x - CONV_FIRST_EPI_WARP * 32;      // 0..255
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // H rows are natural (img, y, x); 1x7 rows are (x, img, y); 7x1 rows are (y, img, x)
    const int rho1 = (r & 7) * 16 + (r >> 6) * 8 + ((r >> 3) & 7);          // natural row r  -> R1 row (x-major)
    const int rho2 = (r & 7) * 16 + ((r >> 3) & 1) * 8 + (r >> 4);          // 1x7 row r      -> R2 row (y-major)
    const int rnat = ((r >> 3) & 1) * 64 + (r >> 4) * 8 + (r & 7);          // 7x1 row r      -> natural row
    const uint32_t s_bias_up = sbase + B17_BIAS_UP;                          // [896] fp32: the `up` bias of the current block
    if (p.pdl) pdl_wait();                                      // the residual of the first block is the previous layer's output
    int blk = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int m0 = tile * CONV_BM;
      for (int j = 0; j < p.n_blocks; ++j, ++blk) {
        const uint32_t bpar = blk & 1;
        const float* bias = p.bias + static_cast<size_t>(j) * B17_BIAS_PER_BLOCK;
        const float4* bias_h = reinterpret_cast<const float4*>(bias + h * 64);   // this warp's 64 columns of each 128-column part
        float4 bq[16];
        // ---- H: columns [0,128) -> R1 (b1a), then [128,256) -> R0 (b0); this warp's half = K-block h of either.
        // The 64 bias values of a part are fetched BEFORE the wait that precedes it, so their latency is never exposed.
#pragma unroll
        for (int e = 0; e < 16; ++e) bq[e] = __ldg(bias_h + e);
        mbar_wait(accH_full, bpar, 45);
        tc_fence_after();
        if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 8);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          const uint32_t row_addr = part == 0 ? r1k(h) + static_cast<uint32_t>(rho1 * 128) : r0a + static_cast<uint32_t>(h) * B17_UNIT + static_cast<uint32_t>(r * 128);
          const uint32_t swz = part == 0 ? (rho1 & 7) : (r & 7);
          if (part == 0) b17_epi_row64_rot(tq + static_cast<uint32_t>(h * 64), bq, row_addr, swz, lane);
          else b17_epi_row64(tq + static_cast<uint32_t>(128 + h * 64), bq, row_addr, swz);
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(part == 0 ? r1_ready : r0_ready);
          if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 9 + part);
#pragma unroll
          for (int e = 0; e < 16; ++e) bq[e] = __ldg(bias_h + (part + 1) * 32 + e);      // next part / the 1x7 conv
        }
        // ---- 1x7 -> R2 (b1b), 7x1 -> R3 (b1c, natural rows, in the data rows of R1)
#pragma unroll
        for (int conv = 0; conv < 2; ++conv) {
          mbar_wait(conv == 0 ? acc17_full : acc71_full, bpar, 46);
          tc_fence_after();
          if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 11 + 2 * conv);
          const uint32_t row_addr = conv == 0 ? r2k(h) + static_cast<uint32_t>(rho2 * 128) : r1k(h) + static_cast<uint32_t>(rnat * 128);
          const uint32_t swz = conv == 0 ? (rho2 & 7) : (rnat & 7);
          if (conv == 0) b17_epi_row64_rot(tq + static_cast<uint32_t>(256 + h * 64), bq, row_addr, swz, lane);
          else b17_epi_row64(tq + static_cast<uint32_t>(384 + h * 64), bq, row_addr, swz);
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(conv == 0 ? r2_ready : r3_ready);
          if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 12 + 2 * conv);
          if (conv == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) bq[e] = __ldg(bias_h + 96 + e);
            // every epilogue warp is past the previous block's `up` (acc17_full implies all eight r1_ready arrivals): refill its bias table
            if (et < 224) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(bias + 512) + et);
              sts128(s_bias_up + et * 16, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
            }
          }
        }
        named_bar_sync(1, CONV_EPI_WARPS * 32);                 // bias table visible to all epilogue warps
        // ---- up: + bias + x, ReLU, fp16 -> y, in 14 groups of 64 columns (128 bytes per row); this warp takes the groups
        // g = h (mod 2) of its 32 rows.  Global memory is only touched with whole 128-byte rows per 8 lanes (4 rows per
        // instruction): "thread = row" accesses cost one LSU wavefront per lane and were 8x slower.  A 4 KB warp-private
        // scratch (R2's data rows, idle during `up`) transposes between the two mappings:
        //   residual: LDG (coalesced, one group ahead, in registers) -> scratch -> own row -> + acc + bias, ReLU, fp16 ->
        //   scratch (in place) -> coalesced mapping -> STG.
        {
          const int ew = warp - CONV_FIRST_EPI_WARP;                                    // 0..7
          const uint32_t sc = r2k(ew >> 2) + static_cast<uint32_t>((ew & 3) * 4096);
          const int lr0 = lane >> 3, pc = lane & 7;                                     // coalesced mapping: row 4 i + lr0, 16-byte piece pc
          const int grow0 = m0 + quarter * 32;
          const __half* xg = p.xptr[j] + static_cast<size_t>(grow0 + lr0) * B17_C + pc * 8;
          __half* yg = const_cast<__half*>(p.xptr[j + 1]) + static_cast<size_t>(grow0 + lr0) * B17_C + pc * 8;
          const uint32_t sc_co = sc + static_cast<uint32_t>(lr0 * 128);                  // + i * 512, chunk pc ^ ((4 i + lr0) & 7)
          const uint32_t sc_own = sc + static_cast<uint32_t>(lane * 128);
          const uint32_t own_swz = lane & 7;
          uint4 rp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rp[i] = (grow0 + 4 * i + lr0 < p.M_total && !(p.dbg & B17_DBG_MASK & 1)) ? ld_cg_v4(xg + static_cast<size_t>(4 * i) * B17_C + h * 64) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int gi = 0; gi < 7; ++gi) {
            const int g = 2 * gi + h;                           // 64-column group; N tile g >> 2 (tile 3 has groups 12, 13 only)
            const int t = g >> 2;
            __syncwarp();                                       // the previous group's coalesced reads of the scratch are done
#pragma unroll
            for (int i = 0; i < 8; ++i) sts128(sc_co + static_cast<uint32_t>(i * 512 + ((pc ^ ((4 * i + lr0) & 7)) << 4)), rp[i]);
            if (gi + 1 < 7) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rp[i] = (grow0 + 4 * i + lr0 < p.M_total && !(p.dbg & B17_DBG_MASK & 1)) ? ld_cg_v4(xg + static_cast<size_t>(4 * i) * B17_C + (g + 2) * 64) : make_uint4(0u, 0u, 0u, 0u);
            }
            if (gi == 0 || ((g & 3) < 2)) {                     // first group of this warp in tile t
              mbar_wait(&accU_full[t & 1], (t >> 1) & 1, 47);
              tc_fence_after();
              if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 15 + t);
            }
            __syncwarp();                                       // residual rows are in the scratch
            {
              const uint32_t ta = tq + static_cast<uint32_t>((t & 1) * 256 + (g & 3) * 64);
              const uint32_t ba = s_bias_up + static_cast<uint32_t>(g * 256);
              uint32_t a[16], b[16];
              auto chunk = [&](uint32_t (&acc)[16], int c) {     // residual + bias, ReLU, fp16, in place in the scratch row
                const uint32_t a_lo = sc_own + (((2 * c) ^ own_swz) << 4), a_hi = sc_own + (((2 * c + 1) ^ own_swz) << 4);
                const uint4 r_lo = lds128(a_lo), r_hi = lds128(a_hi);
                const uint4 q0 = lds128(ba + c * 64), q1 = lds128(ba + c * 64 + 16), q2 = lds128(ba + c * 64 + 32), q3 = lds128(ba + c * 64 + 48);
                b17_add_res(acc, r_lo, r_hi);
                uint4 lo, hi;
                b17_pack_chunk(acc, as_f4(q0), as_f4(q1), as_f4(q2), as_f4(q3), lo, hi);
                sts128(a_lo, lo);
                sts128(a_hi, hi);
              };
              tmem_ld_32x16(ta, a);
              tmem_ld_wait(a);
              tmem_ld_32x16(ta + 16, b);
              chunk(a, 0);
              tmem_ld_wait(b);
              tmem_ld_32x16(ta + 32, a);
              chunk(b, 1);
              tmem_ld_wait(a);
              tmem_ld_32x16(ta + 48, b);
              chunk(a, 2);
              tmem_ld_wait(b);
              chunk(b, 3);
            }
            const bool last = (g & 3) >= 2 || g >= 12;          // last group of this warp in tile t: the accumulator is read
            if (last) tc_fence_before();
            __syncwarp();
            if (last && lane == 0) mbar_arrive(&accU_empty[t & 1]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = lds128(sc_co + static_cast<uint32_t>(i * 512 + ((pc ^ ((4 * i + lr0) & 7)) << 4)));
              if (grow0 + 4 * i + lr0 < p.M_total && !(p.dbg & B17_DBG_MASK & 2)) st_global_v4(yg + static_cast<size_t>(4 * i) * B17_C + g * 64, v);
            }
            if (last && t < 3) {                                // this warp's share of N tile t is on its way to y
              fence_proxy_async_all();
              __syncwarp();
              if (lane == 0) mbar_arrive(&y_part[t]);
            }
          }
        }
        if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 19);
        fence_proxy_async_all();                                // generic-proxy writes of y -> the TMA loads of the next block's H
        __syncwarp();
        if (lane == 0) mbar_arrive(y_done);
        if (warp == CONV_FIRST_EPI_WARP) B17_TRACE(j, 20);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, 512u);
  }
}

}  // namespace fire
