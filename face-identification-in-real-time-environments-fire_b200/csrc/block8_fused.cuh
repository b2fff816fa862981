// block8_fused.cuh - the tail of one Block8 (Inception-ResNet-C) block of FaceNet as ONE tcgen05 launch:
// 1x3 conv -> 3x1 conv -> `up` conv + bias + residual (+ ReLU), i.e. three of the block's four launches.
//
// Reference graph (executed by onnxruntime at facenet_gpu.py:127), per block, on a 3 x 3 x 1792 map:
//     b0  = relu(bn(conv1x1 1792->192 (x)))           b1a = relu(bn(conv1x1 1792->192 (x)))      <- "heads": conv_igemm_kernel
//     b1b = relu(bn(conv1x3 192->192 (b1a)))          b1c = relu(bn(conv3x1 192->192 (b1b)))     <- this kernel
//     y   = [relu](x + s * (conv1x1 384->1792 ([b0 | b1c]) + bias))                              <- this kernel
//
// Why this split.  With 256 images the stage has only M = 2304 rows: layer by layer, every launch pays ~5 us of launch
// dependency + first operand + drain for 2.5-7 us of MMAs.  The heads GEMM (K = 1792) needs every channel of x
// and must stay a grid-wide step; everything after it only needs the heads output of the SAME images.  So a CTA owns
// (a group of 13 images) x (256 of the 1792 `up` output channels): it recomputes the small 1x3 / 3x1 chain of its image
// group redundantly (the seven CTAs of a group all do: 2 x 0.25 GFLOP against 1.3 us of launch overhead each) and then
// produces its own N tile of y.  20 groups x 7 N tiles = 140 CTAs, one wave, no inter-CTA dependency.
//
// Row orders (13 images x 9 positions = 117 rows of one UMMA M tile):
//   1x3  A operand = TMA box (64 ch, 3 x, 3 y, 13 img) of the heads output read at x0 = s - 1: the zero padding is the
//        TMA's out-of-bounds fill, rows are natural (img, y, x).  D1[128 x 192], epilogue -> R2 in y-MAJOR order.
//   3x1  R2 keeps b1b as row = 40 y + 3 img + x (40-row slots, 48 zero rows on either side): tap r of the 3 x 1 conv for ALL
//        outputs is the same buffer shifted by 40 (r - 1) rows, shifts past the image edge read zero rows (block17_fused's
//        trick).  D2[128 x 192] in y-major order, epilogue -> R3 (in place over R2's data rows).
//   up   A = [b0 | R3], b0 loaded y-major by three TMA boxes (64 ch, 3 x, 1 y, 13 img) per K-block at 40-row slots;
//        the residual x is "one more K block" x identity (conv_igemm's trick), y-major the same way; the epilogue adds
//        the bias, converts and stages y-major tiles that three TMA stores per 64 columns write back.
//
// Everything that streams (weights as pre-swizzled byte images, activation boxes, residual boxes) goes through ONE ring
// of eight 16 KB slots in consumption order: 64 units per CTA, dealt to four issuing warps (unit u -> issuer u % 4, slot
// u % 8: a slot always has the same owner, parity waits cannot alias).  Weight units do not depend on the previous
// launch and are requested before griddepcontrol.wait.
#pragma once

#include "block17_fused.cuh"

namespace fire {

constexpr int B8_C = 1792;
constexpr int B8_MID = 192;
constexpr int B8_IMGS = 13;                         // images per CTA group: 117 rows of the 128-row M tile
constexpr int B8_NT = 256;                          // `up` output channels per CTA
constexpr int B8_NTILES = B8_C / B8_NT;             // 7
constexpr int B8_UNIT = 16384;
constexpr int B8_SLOTS = 8;
constexpr int B8_ISSUERS = 4;
constexpr int B8_UNITS = 64;
constexpr int B8_WMID_BYTES = B8_MID * 64;          // 12288: [192 x 32] SWIZZLE_64B
constexpr int B8_WMID_UNITS = 36;                   // 1x3: 18, 3x1: 18
constexpr int B8_WUP_UNITS = 12;                    // per N tile: [256 x 32] SWIZZLE_64B x 12
constexpr int B8_BOX_NAT = 117 * 128;               // bytes of a (64, 3, 3, 13) box
constexpr int B8_BOX_YM = 39 * 128;                 // bytes of a (64, 3, 1, 13) box
constexpr int B8_SLOT_ROWS = 40;                    // y-major slot: 39 rows used, 1024-byte aligned
constexpr int B8_PAD_ROWS = 48, B8_DATA_ROWS = 120;
constexpr int B8_THREADS = 32 * 13;                 // warps: 0, 10, 11, 12 issuers; 1 MMA; 2-9 epilogue
constexpr int B8_BIAS_PER_BLOCK = 2 * B8_MID + B8_C;
constexpr int B8_MAX_BLOCKS = 8;
constexpr int B8_TRACE_SLOTS = 16;

constexpr uint32_t B8_R2 = 0;                                                   // [pad][data k0][pad][data k1][pad][data k2][pad]
constexpr uint32_t B8_R2_BYTES = (4 * B8_PAD_ROWS + 3 * B8_DATA_ROWS) * 128;    // 70656
constexpr uint32_t B8_RING = B8_R2 + B8_R2_BYTES;
constexpr uint32_t B8_IDENT = B8_RING + B8_SLOTS * B8_UNIT;                     // 64 x 64 identity, SWIZZLE_128B
constexpr uint32_t B8_BIAS_MID = B8_IDENT + 64 * 128;                           // [384] fp32
constexpr uint32_t B8_BIAS_UP = B8_BIAS_MID + 2 * B8_MID * 4;                   // [256] fp32
constexpr uint32_t B8_BARS = B8_BIAS_UP + B8_NT * 4;
constexpr uint32_t B8_SMEM = B8_BARS + 256 + 1024;                              // + alignment slack
static_assert(B8_R2_BYTES % 1024 == 0 && B8_RING % 1024 == 0 && B8_IDENT % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(B8_SMEM <= 232448, "shared memory budget");
__host__ __device__ constexpr uint32_t b8_data(int k) { return B8_R2 + static_cast<uint32_t>((B8_PAD_ROWS + k * (B8_DATA_ROWS + B8_PAD_ROWS)) * 128); }

struct B8Params {
  CUtensorMap hmap_nat;     // heads output X [B][3][3][576]: box (64, 3, 3, 13), SWIZZLE_128B
  CUtensorMap hmap_ym;      // same buffer: box (64, 3, 1, 13)
  CUtensorMap xmap_ym;      // block input x [B][3][3][1792] (the residual): box (64, 3, 1, 13)
  CUtensorMap ymap_ym;      // block output y: box (64, 3, 1, 13)
  const uint8_t* wmid;      // 36 units of 12288 B
  const uint8_t* wup;       // 7 x 12 units of 16384 B
  const float* bias;        // [192 | 192 | 1792]
  int n_groups, relu, pdl, b1a_coff, b0_coff;
  __half* gap_out;          // last block only: the global average pool of y, [n_images][gap_ld] fp16 (replaces gap_kernel); else nullptr
  int gap_ld, n_images;
  long long* trace;         // optional: [gridDim.x][B8_TRACE_SLOTS] globaltimer stamps
};

// unit u of the stream -> what it is.  kind 0: mid weights (index a), 1: up weights (index a), 2: 1x3 activation box
// (tap a, K-block b), 3: b0 K-block a, 4: residual K-block a
__device__ __forceinline__ void b8_unit(int u, int& kind, int& a, int& b) {
  b = 0;
  if (u < 27) {
    const int t = u / 3, r = u - 3 * t;
    if (r < 2) { kind = 0; a = 2 * t + r; } else { kind = 2; a = t / 3; b = t - 3 * a; }
  } else if (u < 45) {
    kind = 0; a = 18 + (u - 27);
  } else if (u < 54) {
    const int t = (u - 45) / 3, r = (u - 45) - 3 * t;
    if (r < 2) { kind = 1; a = 2 * t + r; } else { kind = 3; a = t; }
  } else if (u < 60) {
    kind = 1; a = 6 + (u - 54);
  } else {
    kind = 4; a = u - 60;
  }
}

#define B8_TRACE(slot_) do { if (p.trace && lane == 0) p.trace[static_cast<size_t>(blockIdx.x) * B8_TRACE_SLOTS + (slot_)] = globaltimer_ns(); } while (0)

// n_chunks 16-column chunks of this thread's accumulator row -> + bias, [ReLU], fp16 -> swizzled 128-byte rows.
// Chunk i covers columns c0 + 16 i; `row_of(slice)` gives the shared-memory address of this thread's row in the
// 64-column slice the chunk belongs to.  The TMEM load of chunk i + 1 is in flight while chunk i is converted.
template <int kChunks, typename RowOf>
__device__ __forceinline__ void b8_epi_chunks(uint32_t taddr, int c0, uint32_t s_bias, bool relu, bool valid, uint32_t swz, RowOf row_of) {
  uint32_t buf[2][16];
  __syncwarp();
  tmem_ld_32x16(taddr + static_cast<uint32_t>(c0), buf[0]);
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    uint32_t (&r)[16] = buf[i & 1];
    tmem_ld_wait(r);
    if (i + 1 < kChunks) tmem_ld_32x16(taddr + static_cast<uint32_t>(c0 + 16 * (i + 1)), buf[(i + 1) & 1]);
    const int c = c0 + 16 * i;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 b = as_f4(lds128(s_bias + static_cast<uint32_t>((c + 4 * e) * 4)));
      r[4 * e] = __float_as_uint(__uint_as_float(r[4 * e]) + b.x);
      r[4 * e + 1] = __float_as_uint(__uint_as_float(r[4 * e + 1]) + b.y);
      r[4 * e + 2] = __float_as_uint(__uint_as_float(r[4 * e + 2]) + b.z);
      r[4 * e + 3] = __float_as_uint(__uint_as_float(r[4 * e + 3]) + b.w);
    }
    if (valid) conv_stage_chunk(r, relu, row_of(c >> 6), static_cast<uint32_t>((c & 63) >> 3), swz);
  }
}

__global__ void __launch_bounds__(B8_THREADS, 1)
block8_fused_kernel(const __grid_constant__ B8Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B8_BARS);
  uint64_t* full = bars;                          // [8]
  uint64_t* empty = full + B8_SLOTS;              // [8]
  uint64_t* acc1_full = empty + B8_SLOTS;
  uint64_t* acc2_full = acc1_full + 1;
  uint64_t* accU_full = acc2_full + 1;
  uint64_t* r2_ready = accU_full + 1;
  uint64_t* r3_ready = r2_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r3_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = static_cast<int>(blockIdx.x) / B8_NTILES, nt = static_cast<int>(blockIdx.x) - group * B8_NTILES;
  const int img0 = group * B8_IMGS;
  if (warp == 0) B8_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.hmap_nat); tma_prefetch_desc(&p.hmap_ym); tma_prefetch_desc(&p.xmap_ym); tma_prefetch_desc(&p.ymap_ym);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < B8_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(acc1_full, 1); mbar_init(acc2_full, 1); mbar_init(accU_full, 1);
      mbar_init(r2_ready, CONV_EPI_WARPS); mbar_init(r3_ready, CONV_EPI_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, 512u);
  }
  {
    // R2 (pads and the never-written slot rows) = 0; identity; bias tables.  None of it depends on the previous launch.
    for (int i = threadIdx.x; i < static_cast<int>(B8_R2_BYTES / 16); i += B8_THREADS)
      reinterpret_cast<uint4*>(smem + B8_R2)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int t = threadIdx.x; t < 512; t += B8_THREADS) {          // 64 rows x 8 units (conv_igemm.cuh)
      const int row = t >> 3, u = t & 7;
      const int e = row - u * 8;
      const uint32_t one = (e & 1) ? 0x3C000000u : 0x00003C00u;
      const int wi = (e >= 0 && e < 8) ? (e >> 1) : -1;
      *reinterpret_cast<uint4*>(smem + B8_IDENT + sw128_offset(row, u)) =
          make_uint4(wi == 0 ? one : 0u, wi == 1 ? one : 0u, wi == 2 ? one : 0u, wi == 3 ? one : 0u);
    }
    for (int i = threadIdx.x; i < 2 * B8_MID + B8_NT; i += B8_THREADS) {
      const float v = i < 2 * B8_MID ? __ldg(p.bias + i) : __ldg(p.bias + 2 * B8_MID + nt * B8_NT + (i - 2 * B8_MID));
      reinterpret_cast<float*>(smem + B8_BIAS_MID)[i] = v;            // the two tables are contiguous
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) pdl_launch_dependents();

  const uint32_t ring = sbase + B8_RING;
  const int issuer = warp == 0 ? 0 : (warp >= 10 ? warp - 9 : -1);
  if (issuer >= 0) {
    // ---------------------------------------------------------------- the unit stream
    bool waited = !p.pdl;
    for (int u = issuer; u < B8_UNITS; u += B8_ISSUERS) {
      const int slot = u & (B8_SLOTS - 1);
      mbar_wait(&empty[slot], ((static_cast<uint32_t>(u) >> 3) & 1u) ^ 1u, 60);
      int kind, a, b;
      b8_unit(u, kind, a, b);
      if (kind >= 2 && !waited) { pdl_wait(); waited = true; }       // activations come from the previous launches
      if (elect_one()) {
        uint8_t* dst = smem + B8_RING + static_cast<uint32_t>(slot) * B8_UNIT;
        const uint32_t dst32 = ring + static_cast<uint32_t>(slot) * B8_UNIT;
        if (kind == 0) {
          mbar_arrive_expect_tx(&full[slot], B8_WMID_BYTES);
          bulk_copy_g2s(dst32, p.wmid + static_cast<size_t>(a) * B8_WMID_BYTES, B8_WMID_BYTES, &full[slot]);
        } else if (kind == 1) {
          mbar_arrive_expect_tx(&full[slot], B8_UNIT);
          bulk_copy_g2s(dst32, p.wup + (static_cast<size_t>(nt) * B8_WUP_UNITS + a) * B8_UNIT, B8_UNIT, &full[slot]);
        } else if (kind == 2) {
          mbar_arrive_expect_tx(&full[slot], B8_BOX_NAT);
          tma_load_4d(dst, &p.hmap_nat, &full[slot], p.b1a_coff + b * 64, a - 1, 0, img0);
        } else {
          const CUtensorMap* m = kind == 3 ? &p.hmap_ym : &p.xmap_ym;
          const int c = kind == 3 ? p.b0_coff + a * 64 : nt * B8_NT + a * 64;
          mbar_arrive_expect_tx(&full[slot], 3 * B8_BOX_YM);
#pragma unroll
          for (int y = 0; y < 3; ++y) tma_load_4d(dst + y * B8_SLOT_ROWS * 128, m, &full[slot], c, 0, y, img0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc192 = umma_idesc_f16(CONV_BM, B8_MID), idesc256 = umma_idesc_f16(CONV_BM, B8_NT), idesc64 = umma_idesc_f16(CONV_BM, 64);
    auto slot_addr = [ring](int u) { return ring + static_cast<uint32_t>(u & (B8_SLOTS - 1)) * B8_UNIT; };
    auto wait_unit = [&](int u, int tag) { mbar_wait(&full[u & (B8_SLOTS - 1)], (static_cast<uint32_t>(u) >> 3) & 1u, tag); };
    int u = 0;
    // ---- 1x3: D1 = sum over taps s, K-blocks kb of box(s, kb) * W^T
    for (int t = 0; t < 9; ++t, u += 3) {
      wait_unit(u, 61); wait_unit(u + 1, 62); wait_unit(u + 2, 63);
      tc_fence_after();
      if (t == 0) B8_TRACE(1);
      if (elect_one()) {
        const uint32_t a0 = slot_addr(u + 2);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base, umma_desc_sw128(a0 + kk * 32), umma_desc_swz(slot_addr(u + (kk >> 1)) + (kk & 1) * 32, 64), idesc192, (t | kk) != 0 ? 1u : 0u);
        umma_commit(&empty[u & (B8_SLOTS - 1)]);
        umma_commit(&empty[(u + 1) & (B8_SLOTS - 1)]);
        umma_commit(&empty[(u + 2) & (B8_SLOTS - 1)]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc1_full);
    __syncwarp();
    B8_TRACE(2);
    // ---- 3x1: D2 = sum over taps r of R2[rows + 40 (r - 1)] * W^T
    mbar_wait(r2_ready, 0, 64);
    tc_fence_after();
    B8_TRACE(3);
    for (int t = 0; t < 9; ++t, u += 2) {
      const int r = t / 3, kb = t - 3 * r;
      wait_unit(u, 65); wait_unit(u + 1, 66);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = sbase + b8_data(kb) + static_cast<uint32_t>((r - 1) * B8_SLOT_ROWS * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + 256u, umma_desc_sw128(a0 + kk * 32), umma_desc_swz(slot_addr(u + (kk >> 1)) + (kk & 1) * 32, 64), idesc192, (t | kk) != 0 ? 1u : 0u);
        umma_commit(&empty[u & (B8_SLOTS - 1)]);
        umma_commit(&empty[(u + 1) & (B8_SLOTS - 1)]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc2_full);
    __syncwarp();
    B8_TRACE(4);
    // ---- up: DU = [b0 | R3] * Wu^T (this CTA's 256 columns), then + x * I
    mbar_wait(r3_ready, 0, 67);
    tc_fence_after();
    B8_TRACE(5);
    for (int kb = 0; kb < 6; ++kb) {
      wait_unit(u, 68); wait_unit(u + 1, 69);
      if (kb < 3) wait_unit(u + 2, 70);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = kb < 3 ? slot_addr(u + 2) : sbase + b8_data(kb - 3);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base, umma_desc_sw128(a0 + kk * 32), umma_desc_swz(slot_addr(u + (kk >> 1)) + (kk & 1) * 32, 64), idesc256, (kb | kk) != 0 ? 1u : 0u);
        umma_commit(&empty[u & (B8_SLOTS - 1)]);
        umma_commit(&empty[(u + 1) & (B8_SLOTS - 1)]);
        if (kb < 3) umma_commit(&empty[(u + 2) & (B8_SLOTS - 1)]);
      }
      __syncwarp();
      u += kb < 3 ? 3 : 2;
    }
    for (int q = 0; q < 4; ++q, ++u) {
      wait_unit(u, 71);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = slot_addr(u), i0 = sbase + B8_IDENT;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + static_cast<uint32_t>(q * 64), umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(i0 + kk * 32), idesc64, 1u);
        umma_commit(&empty[u & (B8_SLOTS - 1)]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accU_full);
    __syncwarp();
    B8_TRACE(6);
  } else if (warp >= CONV_FIRST_EPI_WARP && warp < CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue (8 warps: TMEM lane quarter x column half)
    const int quarter = warp & 3, h = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const int m = quarter * 32 + lane;                          // accumulator row of this thread
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_bias_mid = sbase + B8_BIAS_MID, s_bias_up = sbase + B8_BIAS_UP;
    // ---- 1x3: natural row m = 9 img + 3 y + x  ->  R2 row 40 y + 3 img + x
    {
      const int img = m / 9, rem = m - 9 * img, yy = rem / 3, xx = rem - 3 * yy;
      const int rho2 = B8_SLOT_ROWS * yy + 3 * img + xx;
      mbar_wait(acc1_full, 0, 72);
      tc_fence_after();
      if (warp == CONV_FIRST_EPI_WARP) B8_TRACE(8);
      b8_epi_chunks<6>(tq, h * 96, s_bias_mid, true, m < 9 * B8_IMGS, static_cast<uint32_t>(rho2 & 7),
                       [&](int k) { return sbase + b8_data(k) + static_cast<uint32_t>(rho2 * 128); });
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(r2_ready);
      if (warp == CONV_FIRST_EPI_WARP) B8_TRACE(9);
    }
    // ---- 3x1: y-major row m -> R3 row m (in place over R2's data rows: every 3x1 MMA has completed)
    {
      mbar_wait(acc2_full, 0, 73);
      tc_fence_after();
      if (warp == CONV_FIRST_EPI_WARP) B8_TRACE(10);
      b8_epi_chunks<6>(tq + 256u, h * 96, s_bias_mid + B8_MID * 4, true, true, static_cast<uint32_t>(m & 7),
                       [&](int k) { return sbase + b8_data(k) + static_cast<uint32_t>(m * 128); });
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(r3_ready);
      if (warp == CONV_FIRST_EPI_WARP) B8_TRACE(11);
    }
    // ---- up: + bias, [ReLU], fp16 -> four y-major staging tiles of 64 columns (R2's data regions and ring slot 0 are free
    // once every MMA has completed) -> TMA stores
    {
      mbar_wait(accU_full, 0, 74);
      tc_fence_after();
      if (warp == CONV_FIRST_EPI_WARP) B8_TRACE(12);
      b8_epi_chunks<8>(tq, h * 128, s_bias_up, p.relu != 0, true, static_cast<uint32_t>(m & 7),
                       [&](int g) { return (g < 3 ? sbase + b8_data(g) : ring) + static_cast<uint32_t>(m * 128); });
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, CONV_EPI_WARPS * 32);
      if (p.gap_out) {
        // global average pool of this CTA's 13 images x 256 channels, read back from the staged fp16 tiles: the same values
        // gap_kernel would read from y, summed in the same order (p = 3 y + x) in fp32, scaled by 1/9, rounded once
        for (int w = threadIdx.x - CONV_FIRST_EPI_WARP * 32; w < B8_IMGS * 32; w += CONV_EPI_WARPS * 32) {
          const int img = w >> 5, pc = w & 31, g = pc >> 3, u = pc & 7;
          if (img0 + img >= p.n_images) continue;
          const uint32_t sg = g < 3 ? sbase + b8_data(g) : ring;
          float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int y = 0; y < 3; ++y)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
              const int row = B8_SLOT_ROWS * y + 3 * img + x;
              const uint4 v = lds128(sg + static_cast<uint32_t>(row * 128 + ((u ^ (row & 7)) << 4)));
              const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __half2 hh;
                memcpy(&hh, &wv[e], 4);
                const float2 f = __half22float2(hh);
                s[2 * e] += f.x; s[2 * e + 1] += f.y;
              }
            }
          const float inv = 1.0f / 9.0f;
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 hh = __floats2half2_rn(s[2 * e] * inv, s[2 * e + 1] * inv);
            memcpy(&o[e], &hh, 4);
          }
          st_global_v4(p.gap_out + static_cast<size_t>(img0 + img) * p.gap_ld + nt * B8_NT + g * 64 + u * 8, make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      if (warp == CONV_FIRST_EPI_WARP) {
        B8_TRACE(13);
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int y = 0; y < 3; ++y)
              tma_store_4d(&p.ymap_ym, (g < 3 ? sbase + b8_data(g) : ring) + static_cast<uint32_t>(y * B8_SLOT_ROWS * 128), nt * B8_NT + g * 64, 0, y, img0);
          bulk_commit_group();
          bulk_wait_all();                                      // stores complete before the CTA exits
        }
        __syncwarp();
        B8_TRACE(14);
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
