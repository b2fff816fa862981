// block35_fused.cuh - the five Block35 (Inception-ResNet-A) blocks of FaceNet as ONE persistent tcgen05 kernel.
//
// Reference graph (executed by onnxruntime at facenet_gpu.py:127), per block, on a 17 x 17 x 256 map:
//     b0 = cbr1x1 256->32 (x)      b1 = cbr3x3 32->32 (cbr1x1 256->32 (x))      b2 = cbr3x3 (cbr3x3 (cbr1x1 256->32 (x)))
//     y  = relu(x + 0.17 * (conv1x1 96->256 ([b0 | b1 | b2]) + bias))                        (cbr = conv + BN + ReLU, 'same')
// The engine's plan runs this as four launches per block (heads 256->96, the two first 3x3 as one block-diagonal 64->64
// conv, the second 3x3 of branch 2, up): 20 launches of ~14 us for 70 GFLOP.  No convolution mixes images, so ONE CTA
// can take an image through a whole block with every intermediate in shared memory (same idea as block17_fused.cuh); only
// the block's input x and output y live in memory.  Which CTA runs which (block, image) is a scheduling question: see
// "Work order" below (round 2: balanced over the SMs, an image's chain wanders from CTA to CTA through flags).
//
// Geometry.  The 3 x 3 'same' convs run on a PITCHED, zero-bordered copy of the image: position (y, x) lives in row
// q = (y + 1) * 18 + (x + 1) of region P12 (pitch 18 = 17 + one shared zero column; one zero row above and below).
// With outputs indexed m = y * 18 + x (m < 306; x = 17 is a dead column), the input of tap (r, s) for ALL outputs is
// the same buffer shifted by r * 18 + s rows: each tap is a plain K = 64 (or 32) MMA on a row-shifted window, in three
// M tiles of 128 rows, with no im2col and no boundary code.  Dead rows are computed and never stored.
//   H     three natural-order M tiles of x (plain 2-D TMA, 289 rows): D[mt] = x * Wh^T (N = 96 = [b1a | b2a | b0]);
//         epilogue: b1a | b2a -> P12 row q (64 channels), b0 -> AU0 row m, channels 0-31
//   conv1 block-diagonal 3x3 64 -> 64 = [b1b | b2b]: b1b -> AU0 row m, channels 32-63; b2b -> P12 row q, channels 0-31
//         (over b1a, once every conv1 MMA has completed)
//   conv2 3x3 32 -> 32 on P12 channels 0-31: b2c -> AU1 row m
//   up    per M tile: D[128 x 256] = [AU0 | AU1] * Wu^T (K = 96); epilogue + bias + x, ReLU -> y (natural rows), through
//         a warp-private scratch so that global memory only sees whole 64-byte row pieces
// One ring of six 16 KB slots carries everything that streams: the x tiles of H (TMA) and all weights (cp.async.bulk
// from a host-packed stream of swizzled operand images, in consumption order), issued by two warps (unit u by issuer
// u % 2).  TMEM: accumulators of tile mt at columns 128 mt (+96 for conv2); `up` alternates [0,256) / [256,512).
#pragma once

#include <type_traits>

#include "block17_fused.cuh"

namespace fire {

constexpr int B35_C = 256;
constexpr int B35_HW = 17, B35_P = 18;
constexpr int B35_POS = 289;                // positions per image
constexpr int B35_MROWS = 306;              // rows m = y * 18 + x that can hold a position
constexpr int B35_UNIT = 16384;
constexpr int B35_SLOTS = 6;
constexpr int B35_UNITS_PER_BLOCK = 16 + 9 + 2 + 12;      // 39 ring units: H (4 x (W, x, x, x)), conv1 taps, conv2 (2), up (3 x 4)
constexpr int B35_WSLOTS_PER_BLOCK = 4 + 9 + 2 + 4;        // 19 weight units of 16 KB in the stream
constexpr int B35_BIAS_PER_BLOCK = 96 + 64 + 32 + 256;     // 448
constexpr int B35_MAX_BLOCKS = 5;
constexpr int B35_THREADS = 32 * 12;

constexpr uint32_t B35_P12 = 0;                               // 422 rows x 128 B (rows up to 2 * 128 + 38 + 127 are read)
constexpr uint32_t B35_AU0 = 54272;                           // [b0 | b1b]: 306 rows x 128 B (the MMA reads 384; the tail is never used)
constexpr uint32_t B35_AU1 = B35_AU0 + 39936;                 // b2c: 306 rows x 64 B
constexpr uint32_t B35_RING = B35_AU1 + 20480;                // 114688
constexpr uint32_t B35_SCR = B35_RING + B35_SLOTS * B35_UNIT; // 212992: 8 warps x 2 KB
constexpr uint32_t B35_BARS = B35_SCR + 8 * 2048;             // 229376
// Biases of the current block in shared memory (with this much dynamic smem L1 is ~0 KB: a bias __ldg in the epilogue
// is an L2 round trip per chunk - measured 435 cycles per chunk).  [H 96 | conv1 64 | conv2 32] sit in the unused tail
// of AU0's allocation, [up 256] after the barriers.
constexpr uint32_t B35_BIAS_HC = B35_AU0 + B35_MROWS * 128;   // 768 B
constexpr uint32_t B35_BIAS_UP = B35_BARS + 288;              // 1024 B
constexpr uint32_t B35_SMEM = B35_BIAS_UP + 1024 + 1024;      // 231712
static_assert(B35_AU0 % 1024 == 0 && B35_AU1 % 1024 == 0 && B35_RING % 1024 == 0 && B35_SCR % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(B35_SMEM <= 232448, "shared memory budget");

struct B35Params {
  CUtensorMap xmap[B35_MAX_BLOCKS + 1];   // x_0 .. x_n: [B * 289][256] fp16, box 64 columns x 128 rows, SWIZZLE_128B
  const __half* xptr[B35_MAX_BLOCKS + 1];
  const uint8_t* wstream;                 // n_blocks x 19 units of 16 KB
  const float* bias;                      // n_blocks x 448
  int n_blocks, n_images, pdl;
  int* flags;                             // [n_blocks][n_images]: launch number (epoch) of the last launch that wrote block j's y of image i
  int epoch, balance;                     // balance = 0: a CTA takes whole images through all blocks (no flags)
  long long* trace;
};

// Work order.  One TASK = one block of one image (~25 us).  Tasks are numbered block-major, t = j * n_images + img, and CTA w
// takes t = w, w + G, w + 2G, ...: with n_images <= G (or a multiple of G) that is "CTA w owns image w" and a block's input is
// the y this CTA has just written (y_done, CTA-local); otherwise - 256 images on 148 SMs - the chain of an image wanders over
// CTAs and a block's input is published through a per-(block, image) flag (release / acquire at gpu scope).  The predecessor
// task t - n_images always belongs to an EARLIER round of some CTA, CTAs walk their tasks in increasing order and the grid is
// co-resident (the host launches this case cooperatively: the grid only starts when all of it fits at once), so nothing can
// wait in a circle; 1280 tasks then take ceil(1280 / 148) = 9 rounds instead of 2 x 5.
__device__ __forceinline__ int b35_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void b35_wait_flag(const int* f, int epoch, int tag) {
  const long long t0 = clock64();
  while (static_cast<int>(static_cast<uint32_t>(b35_ld_acquire(f)) - static_cast<uint32_t>(epoch)) < 0) {     // wrap-safe: launch numbers only grow
    if (clock64() - t0 > FIRE_WATCHDOG_CYCLES) { printf("fire_b200: block35 flag watchdog (tag %d, block %d)\n", tag, blockIdx.x); __trap(); }
  }
}

// ring unit u of a block: x tile or weight unit, and its size
__device__ __forceinline__ bool b35_is_x(int u) { return u < 16 && (u & 3) != 0; }
__device__ __forceinline__ int b35_wslot(int u) { return u < 16 ? (u >> 2) : u < 27 ? u - 12 : 15 + ((u - 27) & 3); }
__device__ __forceinline__ uint32_t b35_wbytes(int u) {
  return u < 16 ? 12288u : u < 25 ? 8192u : u == 25 ? 10240u : u == 26 ? 8192u : ((u - 27) & 1) ? 8192u : 16384u;
}

// one 16-column chunk: + bias, ReLU, fp16 -> two 16-byte units (the caller places them)
__device__ __forceinline__ void b35_chunk(uint32_t (&acc)[16], uint32_t bias_smem, uint4& lo, uint4& hi) {
  const uint4 q0 = lds128(bias_smem), q1 = lds128(bias_smem + 16), q2 = lds128(bias_smem + 32), q3 = lds128(bias_smem + 48);
  b17_pack_chunk(acc, as_f4(q0), as_f4(q1), as_f4(q2), as_f4(q3), lo, hi);
}

#define B35_TRACE(slot_) do { if (p.trace && lane == 0 && blk < 16) p.trace[(static_cast<size_t>(blockIdx.x) * 16 + blk) * 24 + (slot_)] = globaltimer_ns(); } while (0)

__global__ void __launch_bounds__(B35_THREADS, 1)
block35_fused_kernel(const __grid_constant__ B35Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B35_BARS);
  uint64_t* full = bars;                         // [6]
  uint64_t* empty = full + B35_SLOTS;            // [6]
  uint64_t* accH_full = empty + B35_SLOTS;       // [3]
  uint64_t* acc1_full = accH_full + 3;           // [3]
  uint64_t* acc2_full = acc1_full + 3;           // [3]
  uint64_t* accU_full = acc2_full + 3;           // [2]
  uint64_t* accU_empty = accU_full + 2;          // [2]
  uint64_t* hready = accU_empty + 2;             // [3] H epilogue of tile mt is in P12 / AU0
  uint64_t* b2ready = hready + 3;                // [3] b2b of tile mt is in P12
  uint64_t* aup_ready = b2ready + 3;             // [3] b0, b1b, b2c of tile mt are in AU0 / AU1
  uint64_t* y_done = aup_ready + 3;              // this block's y (= the next block's x) is written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j <= p.n_blocks; ++j) tma_prefetch_desc(&p.xmap[j]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < B35_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }       // a slot is released by two commits
      for (int t = 0; t < 3; ++t) {
        mbar_init(&accH_full[t], 1); mbar_init(&acc1_full[t], 1); mbar_init(&acc2_full[t], 1);
        mbar_init(&hready[t], CONV_EPI_WARPS); mbar_init(&b2ready[t], CONV_EPI_WARPS / 2); mbar_init(&aup_ready[t], CONV_EPI_WARPS);
      }
      for (int b = 0; b < 2; ++b) { mbar_init(&accU_full[b], 2); mbar_init(&accU_empty[b], CONV_EPI_WARPS); }   // one arrival per N half
      mbar_init(y_done, CONV_EPI_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, 512u);
  }
  for (int i = threadIdx.x; i < static_cast<int>(B35_AU0 / 16); i += B35_THREADS)      // P12: zero border (and everything else) once
    reinterpret_cast<uint4*>(smem + B35_P12)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) pdl_launch_dependents();

  const uint32_t p12 = sbase + B35_P12, au0 = sbase + B35_AU0, au1 = sbase + B35_AU1, ring = sbase + B35_RING;
  const int issuer = warp == 0 ? 0 : warp == 10 ? 1 : -1;
  // the previous block of a task's image was this CTA's previous task (CTA-local hand-over) unless the chains wander
  const bool local_chain = !p.balance || (p.n_images % static_cast<int>(gridDim.x)) == 0;
  // k-th task of this CTA -> (block j, image img); false when the CTA has no k-th task.  Balanced: t = w + G k in block-major
  // numbering; otherwise the CTA owns images w, w + G, ... and takes each through all blocks.
  auto task_of = [&](int k, int& j, int& img) {
    if (!local_chain) {
      const int t = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
      j = t / p.n_images; img = t - j * p.n_images;
      return t < p.n_blocks * p.n_images;
    }
    const int q = k / p.n_blocks;
    j = k - q * p.n_blocks; img = static_cast<int>(blockIdx.x) + q * static_cast<int>(gridDim.x);
    return img < p.n_images;
  };

  if (issuer >= 0) {
    // ---------------------------------------------------------------- producers: x tiles (TMA) and the weight stream (bulk copies)
    if (p.pdl) pdl_wait();
    int slot = 0, blk = 0;
    uint32_t ph = 0, gu = 0;
    for (int j, img; task_of(blk, j, img); ++blk) {
      {
        bool gated = j == 0;                                    // x_j of blocks 1.. is the y of block j - 1: this CTA's previous task, or published by flag
        for (int u = 0; u < B35_UNITS_PER_BLOCK; ++u, ++gu) {
          if ((gu & 1) == static_cast<uint32_t>(issuer)) {
            const bool is_x = b35_is_x(u);
            if (is_x && !gated) {
              if (local_chain) mbar_wait(y_done, (blk - 1) & 1, 61);
              else b35_wait_flag(p.flags + static_cast<size_t>(j - 1) * p.n_images + img, p.epoch, 61);
              fence_proxy_async_all();
              gated = true;
            }
            mbar_wait(&empty[slot], ph ^ 1, 62);
            if (elect_one()) {
              if (is_x) {
                mbar_arrive_expect_tx(&full[slot], B35_UNIT);
                tma_load_2d(smem + B35_RING + slot * B35_UNIT, &p.xmap[j], &full[slot], (u >> 2) * 64, img * B35_POS + ((u & 3) - 1) * CONV_BM);
              } else {
                const uint32_t bytes = b35_wbytes(u);
                mbar_arrive_expect_tx(&full[slot], bytes);
                bulk_copy_g2s(ring + static_cast<uint32_t>(slot) * B35_UNIT,
                              p.wstream + (static_cast<size_t>(j) * B35_WSLOTS_PER_BLOCK + b35_wslot(u)) * B35_UNIT, bytes, &full[slot]);
              }
            }
            __syncwarp();
          }
          if (++slot == B35_SLOTS) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 11) {
    // ---------------------------------------------------------------- MMA issuers (two warps)
    // One thread issues a tcgen05.mma every ~110 cycles whatever its size and these GEMMs have N <= 128, so ONE issuer
    // is the bottleneck of H / conv1 / conv2; issuers in different warps overlap (tools/umma_probe.cu part 6).  The
    // three M tiles have independent accumulators: warp 1 owns tiles 0 and 2, warp 11 tile 1 (in `up`: the first / second
    // N half of every tile).  Both walk the same unit
    // sequence and BOTH wait for every fill, also of the units only the other warp reads: mbarrier waits are by phase
    // PARITY, so a warp that skipped a fill of a slot could take the fill before it for the one it is waiting for (the
    // ring wraps every six units; this was an intermittent wrong result), and BOTH commit once on every slot's `empty`
    // barrier (count 2), so neither can fall a whole ring behind the producer either.  `mw`
    // is a compile-time constant of each instantiation (predicates derived from threadIdx are not provably warp-uniform:
    // ptxas would serialise every MMA).
    auto mma_role = [&](auto mw_c) {
    constexpr int mw = decltype(mw_c)::value;
    const uint32_t idesc96 = umma_idesc_f16(CONV_BM, 96), idesc64 = umma_idesc_f16(CONV_BM, 64), idesc32 = umma_idesc_f16(CONV_BM, 32),
                   idesc128 = umma_idesc_f16(CONV_BM, 128);
    int slot = 0, blk = 0;
    uint32_t ph = 0;
    auto next_slot = [&]() { if (++slot == B35_SLOTS) { slot = 0; ph ^= 1; } };
    auto owns = [&](int idx) { return (idx & 1) == mw; };
    // every warp commits exactly once on every unit's `empty` barrier: after its MMAs if it read the slot, at once otherwise
    auto release_slot = [&]() {
      if (elect_one()) umma_commit(&empty[slot]);
      __syncwarp();
    };
    for (int j, img; task_of(blk, j, img); ++blk) {
      {
        const uint32_t bpar = blk & 1;
        // ---- H: D[mt] (TMEM columns 128 mt .. +95) = x[mt] * Wh^T
        if (blk > 0) { mbar_wait(&accU_empty[0], 1, 63); mbar_wait(&accU_empty[1], (blk - 1) & 1, 64); }   // the previous `up` has left TMEM
        tc_fence_after();
        if (mw == 0) B35_TRACE(0);
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&full[slot], ph, 65);
          const int wslot = slot;
          next_slot();
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            mbar_wait(&full[slot], ph, 66);             // BOTH warps observe every fill in order (see above), the owner issues
            if (owns(mt)) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a0 = ring + static_cast<uint32_t>(slot) * B35_UNIT, b0 = ring + static_cast<uint32_t>(wslot) * B35_UNIT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem_base + static_cast<uint32_t>(mt * 128), umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc96, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty[slot]);
                if (kb == 3) umma_commit(&accH_full[mt]);
              }
              __syncwarp();
            } else {
              if (elect_one()) umma_commit(&empty[slot]);
              __syncwarp();
            }
            next_slot();
          }
          if (elect_one()) umma_commit(&empty[wslot]);
          __syncwarp();
        }
        if (mw == 0) B35_TRACE(1);
        // ---- conv1: nine taps, each a K = 64 MMA on P12 shifted by r * 18 + s rows
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&full[slot], ph, 67);
          tc_fence_after();
          const int r = tap / 3, s = tap - 3 * r;
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            if (!owns(mt)) continue;
            if (tap == 0) {                                     // tile mt reads P12 rows q - 19 in [128 mt - 19, 128 mt + 147): all three H epilogues
              for (int t = 0; t < 3; ++t) mbar_wait(&hready[t], bpar, 68);
              tc_fence_after();
              if (mt == 0) B35_TRACE(2);
            }
            if (elect_one()) {
              const uint32_t a0 = p12 + static_cast<uint32_t>((mt * 128 + r * B35_P + s) * 128), b0 = ring + static_cast<uint32_t>(slot) * B35_UNIT;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + static_cast<uint32_t>(mt * 128), umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc64, (tap | k) != 0 ? 1u : 0u);
              if (tap == 8) umma_commit(&acc1_full[mt]);
            }
            __syncwarp();
          }
          release_slot();
          next_slot();
        }
        if (mw == 0) B35_TRACE(3);
        // ---- conv2: nine taps, K = 32, on channels 0-31 of P12 (now b2b); two weight units (taps 0-4, 5-8)
        for (int tap = 0; tap < 9; ++tap) {
          if (tap == 0 || tap == 5) { mbar_wait(&full[slot], ph, 69); tc_fence_after(); }
          const int r = tap / 3, s = tap - 3 * r, tl = tap < 5 ? tap : tap - 5;
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            if (!owns(mt)) continue;
            if (tap == 0) {
              for (int t = 0; t < 3; ++t) mbar_wait(&b2ready[t], bpar, 70);
              tc_fence_after();
              if (mt == 0) B35_TRACE(4);
            }
            if (elect_one()) {
              const uint32_t a0 = p12 + static_cast<uint32_t>((mt * 128 + r * B35_P + s) * 128);
              const uint32_t b0 = ring + static_cast<uint32_t>(slot) * B35_UNIT + static_cast<uint32_t>(tl * 2048);
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_f16(tmem_base + static_cast<uint32_t>(mt * 128 + 96), umma_desc_sw128(a0 + k * 32), umma_desc_swz(b0 + k * 32, 64), idesc32, (tap | k) != 0 ? 1u : 0u);
              if (tap == 8) umma_commit(&acc2_full[mt]);
            }
            __syncwarp();
          }
          if (tap == 4 || tap == 8) {
            release_slot();
            next_slot();
          }
        }
        if (mw == 0) B35_TRACE(5);
        // ---- up: per M tile D[128 x 256] = [AU0 | AU1] * Wu^T in two N halves
        for (int t = 0; t < 3; ++t) mbar_wait(&aup_ready[t], bpar, 71);
        tc_fence_after();
        if (mw == 0) B35_TRACE(6);
        for (int mt = 0; mt < 3; ++mt) {
          if (mt == 2) { mbar_wait(&accU_empty[0], 0, 72); tc_fence_after(); }      // `up` of tile 0 has been read
          const uint32_t d = tmem_base + static_cast<uint32_t>((mt & 1) * 256);
#pragma unroll
          for (int nh = 0; nh < 2; ++nh) {
            mbar_wait(&full[slot], ph, 73);
            if (owns(nh)) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a0 = au0 + static_cast<uint32_t>(mt) * 16384u, b0 = ring + static_cast<uint32_t>(slot) * B35_UNIT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(d + static_cast<uint32_t>(nh * 128), umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc128, k != 0 ? 1u : 0u);
              }
              __syncwarp();
            }
            release_slot();
            next_slot();
            mbar_wait(&full[slot], ph, 74);
            if (owns(nh)) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a0 = au1 + static_cast<uint32_t>(mt) * 8192u, b0 = ring + static_cast<uint32_t>(slot) * B35_UNIT;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  umma_f16(d + static_cast<uint32_t>(nh * 128), umma_desc_swz(a0 + k * 32, 64), umma_desc_swz(b0 + k * 32, 64), idesc128, 1u);
                umma_commit(&accU_full[mt & 1]);                // one of two arrivals per tile (one per N half)
              }
              __syncwarp();
            }
            release_slot();
            next_slot();
          }
        }
        if (mw == 0) B35_TRACE(7);
      }
    }
    };
    if (warp == 1) mma_role(std::integral_constant<int, 0>{});
    else mma_role(std::integral_constant<int, 1>{});
  } else if (warp >= CONV_FIRST_EPI_WARP && warp < CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue (8 warps: TMEM lane quarter x half)
    const int quarter = warp & 3, h = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t scr = sbase + B35_SCR + static_cast<uint32_t>((warp - CONV_FIRST_EPI_WARP) * 2048);
    const int lr0 = lane >> 2, pc = lane & 3;                   // coalesced mapping: row 8 i + lr0, 16-byte piece pc of a 64-byte row
    const int et = threadIdx.x - CONV_FIRST_EPI_WARP * 32;      // 0..255
    const uint32_t bias_hc = sbase + B35_BIAS_HC, bias_up = sbase + B35_BIAS_UP;
    float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (et < B35_BIAS_PER_BLOCK / 4) nb = __ldg(reinterpret_cast<const float4*>(p.bias) + et);
    if (p.pdl) pdl_wait();
    int blk = 0;
    for (int j, img; task_of(blk, j, img); ++blk) {              // (the first task of every CTA is in block 0: nb above)
      {
        const uint32_t bpar = blk & 1;
        // bias table: every warp is done with the previous block's values (first barrier), 112 threads store the float4 they
        // requested one task ago, and request the next task's
        named_bar_sync(1, CONV_EPI_WARPS * 32);
        if (et < B35_BIAS_PER_BLOCK / 4)
          sts128(et < 48 ? bias_hc + et * 16 : bias_up + (et - 48) * 16, make_uint4(__float_as_uint(nb.x), __float_as_uint(nb.y), __float_as_uint(nb.z), __float_as_uint(nb.w)));
        named_bar_sync(1, CONV_EPI_WARPS * 32);
        if (et < B35_BIAS_PER_BLOCK / 4)
          {
            int jn, in_;
            nb = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(task_of(blk + 1, jn, in_) ? jn : 0) * B35_BIAS_PER_BLOCK) + et);
          }
        uint32_t a[16], b[16];
        uint4 lo, hi;
        // ---- H: natural rows.  Chunks 0-3 (b1a | b2a) -> P12 row q, chunks 4, 5 (b0) -> AU0 row m; this half takes 3 chunks
        for (int mt = 0; mt < 3; ++mt) {
          const int pos = mt * 128 + r;
          const bool ok = pos < B35_POS;
          const int y = (pos * 3856) >> 16, x = pos - y * B35_HW;                      // pos / 17
          const int m = y * B35_P + x, q = m + B35_P + 1;
          const uint32_t q_addr = p12 + static_cast<uint32_t>(q * 128), q_swz = q & 7;
          const uint32_t m_addr = au0 + static_cast<uint32_t>(m * 128), m_swz = m & 7;
          auto put = [&](int c) {                               // chunk c of the 96 columns
            if (!ok) return;
            if (c < 4) { sts128(q_addr + (((2 * c) ^ q_swz) << 4), lo); sts128(q_addr + (((2 * c + 1) ^ q_swz) << 4), hi); }
            else { sts128(m_addr + (((2 * (c - 4)) ^ m_swz) << 4), lo); sts128(m_addr + (((2 * (c - 4) + 1) ^ m_swz) << 4), hi); }
          };
          const uint32_t ta = tq + static_cast<uint32_t>(mt * 128 + h * 48);
          const uint32_t bh = bias_hc + static_cast<uint32_t>(h * 48 * 4);
          mbar_wait(&accH_full[mt], bpar, 75);
          tc_fence_after();
          if (warp == CONV_FIRST_EPI_WARP && mt == 0) B35_TRACE(8);
          __syncwarp();
          tmem_ld_32x16(ta, a);
          tmem_ld_wait(a);
          tmem_ld_32x16(ta + 16, b);
          b35_chunk(a, bh, lo, hi); put(3 * h);
          tmem_ld_wait(b);
          tmem_ld_32x16(ta + 32, a);
          b35_chunk(b, bh + 64, lo, hi); put(3 * h + 1);
          tmem_ld_wait(a);
          b35_chunk(a, bh + 128, lo, hi); put(3 * h + 2);
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&hready[mt]);
          if (warp == CONV_FIRST_EPI_WARP && mt == 2) B35_TRACE(9);
        }
        // ---- conv1: half 0 takes b1b (columns 0-31 -> AU0 channels 32-63) tile by tile; half 1 takes b2b (columns 32-63 ->
        // P12 channels 0-31, over b1a) once EVERY conv1 MMA has completed (they all read b1a)
        if (h == 1) { for (int t = 0; t < 3; ++t) mbar_wait(&acc1_full[t], bpar, 76); }     // each commit only covers its own warp's MMAs
        for (int mt = 0; mt < 3; ++mt) {
          const int m = mt * 128 + r;
          const int y = (m * 3641) >> 16, x = m - y * B35_P;                           // m / 18
          const bool ok = h == 0 ? m < B35_MROWS : (x < B35_HW && y < B35_HW);
          const uint32_t row = h == 0 ? au0 + static_cast<uint32_t>(m * 128) : p12 + static_cast<uint32_t>((m + B35_P + 1) * 128);
          const uint32_t swz = h == 0 ? (m & 7) : ((m + B35_P + 1) & 7);
          const uint32_t u0 = h == 0 ? 4u : 0u;
          if (h == 0) mbar_wait(&acc1_full[mt], bpar, 77);
          tc_fence_after();
          if (warp == CONV_FIRST_EPI_WARP && mt == 0) B35_TRACE(10);
          __syncwarp();
          const uint32_t ta = tq + static_cast<uint32_t>(mt * 128 + h * 32);
          tmem_ld_32x16(ta, a);
          tmem_ld_wait(a);
          tmem_ld_32x16(ta + 16, b);
          b35_chunk(a, bias_hc + static_cast<uint32_t>((96 + h * 32) * 4), lo, hi);
          if (ok) { sts128(row + ((u0 ^ swz) << 4), lo); sts128(row + (((u0 + 1) ^ swz) << 4), hi); }
          tmem_ld_wait(b);
          b35_chunk(b, bias_hc + static_cast<uint32_t>((96 + h * 32 + 16) * 4), lo, hi);
          if (ok) { sts128(row + (((u0 + 2) ^ swz) << 4), lo); sts128(row + (((u0 + 3) ^ swz) << 4), hi); }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (h == 1 && lane == 0) mbar_arrive(&b2ready[mt]);
          if (warp == CONV_FIRST_EPI_WARP && mt == 2) B35_TRACE(11);
        }
        // ---- conv2: 32 columns, one chunk per half -> AU1 row m (64-byte rows, SWIZZLE_64B)
        for (int mt = 0; mt < 3; ++mt) {
          const int m = mt * 128 + r;
          mbar_wait(&acc2_full[mt], bpar, 78);
          tc_fence_after();
          if (warp == CONV_FIRST_EPI_WARP && mt == 0) B35_TRACE(12);
          __syncwarp();
          tmem_ld_32x16(tq + static_cast<uint32_t>(mt * 128 + 96 + h * 16), a);
          tmem_ld_wait(a);
          b35_chunk(a, bias_hc + static_cast<uint32_t>((160 + h * 16) * 4), lo, hi);
          if (m < B35_MROWS) {
            const uint32_t row = au1 + static_cast<uint32_t>(m * 64), swz = (m >> 1) & 3;
            sts128(row + (((2 * h) ^ swz) << 4), lo);
            sts128(row + (((2 * h + 1) ^ swz) << 4), hi);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aup_ready[mt]);
          if (warp == CONV_FIRST_EPI_WARP && mt == 2) B35_TRACE(13);
        }
        // ---- up: + bias + x, ReLU -> y.  Per M tile this warp takes the 64-column groups g = h (mod 2), 32 columns at a time:
        // residual rows arrive in the coalesced mapping (4 lanes = one 64-byte row piece), are transposed through the 2 KB
        // scratch to "thread = row", combined, and go back the same way.
        if (!local_chain && j > 0) b35_wait_flag(p.flags + static_cast<size_t>(j - 1) * p.n_images + img, p.epoch, 80);   // the residual x_j came from another CTA
        const __half* xg = p.xptr[j] + static_cast<size_t>(img) * B35_POS * B35_C + pc * 8;
        __half* yg = const_cast<__half*>(p.xptr[j + 1]) + static_cast<size_t>(img) * B35_POS * B35_C + pc * 8;
        const uint32_t sc_own = scr + static_cast<uint32_t>(lane * 64), own_swz = (lane >> 1) & 3;
        const bool prof = p.trace != nullptr && warp == CONV_FIRST_EPI_WARP;
        long long tw[5] = {0, 0, 0, 0, 0};   // cycles: scratch fill + residual request | accumulator wait | first TMEM load | two chunks | write-back + stores
        for (int mt = 0; mt < 3; ++mt) {
          int goff[4];                                          // element offset of the natural row behind scratch row 8 i + lr0, or -1
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = mt * 128 + quarter * 32 + 8 * i + lr0;
            const int y = (m * 3641) >> 16, x = m - y * B35_P;
            goff[i] = (x < B35_HW && y < B35_HW) ? (y * B35_HW + x) * B35_C : -1;
          }
          uint4 rp[4];
          auto load_res = [&](int step) {                       // step = 2 * (group index 0, 1) + 32-column half
            const int col = ((step >> 1) * 2 + h) * 64 + (step & 1) * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) rp[i] = goff[i] >= 0 ? ld_cg_v4(xg + goff[i] + col) : make_uint4(0u, 0u, 0u, 0u);
          };
          load_res(0);
#pragma unroll
          for (int step = 0; step < 4; ++step) {
            const int col = ((step >> 1) * 2 + h) * 64 + (step & 1) * 32;
            long long c0 = prof ? clock64() : 0, c1;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + lr0;
              sts128(scr + static_cast<uint32_t>(row * 64 + ((pc ^ ((row >> 1) & 3)) << 4)), rp[i]);
            }
            if (step + 1 < 4) load_res(step + 1);
            if (prof) { c1 = clock64(); tw[0] += c1 - c0; c0 = c1; }
            if (step == 0) {
              mbar_wait(&accU_full[mt & 1], mt == 1 ? bpar : (mt == 0 ? 0u : 1u), 79);
              tc_fence_after();
              if (warp == CONV_FIRST_EPI_WARP) B35_TRACE(14 + mt);
            }
            __syncwarp();
            if (prof) { c1 = clock64(); tw[1] += c1 - c0; c0 = c1; }
            const uint32_t ta = tq + static_cast<uint32_t>((mt & 1) * 256 + col);
            tmem_ld_32x16(ta, a);
            tmem_ld_wait(a);
            if (prof) { c1 = clock64(); tw[2] += c1 - c0; c0 = c1; }
            tmem_ld_32x16(ta + 16, b);
            {
              const uint32_t a_lo = sc_own + ((0u ^ own_swz) << 4), a_hi = sc_own + ((1u ^ own_swz) << 4);
              b17_add_res(a, lds128(a_lo), lds128(a_hi));
              b35_chunk(a, bias_up + static_cast<uint32_t>(col * 4), lo, hi);
              sts128(a_lo, lo); sts128(a_hi, hi);
            }
            tmem_ld_wait(b);
            {
              const uint32_t a_lo = sc_own + ((2u ^ own_swz) << 4), a_hi = sc_own + ((3u ^ own_swz) << 4);
              b17_add_res(b, lds128(a_lo), lds128(a_hi));
              b35_chunk(b, bias_up + static_cast<uint32_t>((col + 16) * 4), lo, hi);
              sts128(a_lo, lo); sts128(a_hi, hi);
            }
            if (prof) { c1 = clock64(); tw[3] += c1 - c0; c0 = c1; }
            if (step == 3) tc_fence_before();
            __syncwarp();
            if (step == 3 && lane == 0) mbar_arrive(&accU_empty[mt & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + lr0;
              const uint4 v = lds128(scr + static_cast<uint32_t>(row * 64 + ((pc ^ ((row >> 1) & 3)) << 4)));
              if (goff[i] >= 0) st_global_v4(yg + goff[i] + col, v);
            }
            if (prof) { c1 = clock64(); tw[4] += c1 - c0; }
          }
        }
        if (prof && lane == 0 && blk < 16) { for (int k = 0; k < 5; ++k) p.trace[(static_cast<size_t>(blockIdx.x) * 16 + blk) * 24 + 19 + k] = tw[k]; }
        if (warp == CONV_FIRST_EPI_WARP) B35_TRACE(17);
        fence_proxy_async_all();                                // generic-proxy writes of y -> the TMA loads of the next block's H
        __syncwarp();
        if (lane == 0) mbar_arrive(y_done);
        if (!local_chain && warp == CONV_FIRST_EPI_WARP) {     // all eight warps have stored their share of y: publish it
          mbar_wait(y_done, bpar, 81);
          if (lane == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags + static_cast<size_t>(j) * p.n_images + img), "r"(p.epoch) : "memory");
          }
          __syncwarp();
        }
        if (warp == CONV_FIRST_EPI_WARP) B35_TRACE(18);
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
