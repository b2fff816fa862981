// conv_igemm.cuh - persistent implicit-GEMM convolution on tcgen05 (K2 of DESIGN.md).
//
// One launch = one Conv(+folded BN)(+bias)(+residual)(+ReLU) layer of the FaceNet plan
// (reference: the onnxruntime session at facenet_gpu.py:127; graph in SURVEY App. A).
//
//   D[m, n] = bias[n] + sum_k A[m, k] * W[n, k] (+ R[m, n])      m = output pixel (b, ho, wo), n = output channel,
//                                                                 k = (tap r,s ; input channel c), tap-major
//
// EVERYTHING additive runs on the tensor core, so the epilogue is only "convert and store":
//   * bias      : the last MMA of a tile is ones[128 x 16] * biasT[bn x 16]^T, where the bias row holds fp16 hi and lo
//                 parts of the fp32 bias (hi + lo carries ~22 bits);
//   * residual  : x + up(cat) is "one more K block": the residual tile [128 x 64] is TMA-loaded like an activation
//                 K-block and multiplied by a constant 64 x 64 identity (exact: fp16 * 1.0 accumulated in fp32).
//
// Persistent CTAs (one per SM) walk the (m-tile, n-tile) list; warp roles:
//   warp 0      TMA producer: weight K-blocks (always); for 1x1/stride-1 layers also the activation K-blocks (the
//               activation matrix [M, C] is a plain 2-D tensor) and the residual K-blocks;
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=bn_tile, K=16) into one of TWO TMEM accumulators,
//               so the epilogue of tile t overlaps the main loop of tile t+1;
//   warps 2-9   epilogue, two warps per TMEM lane quarter, 16-column chunks interleaved between them:
//               tcgen05.ld -> cvt.rn[.relu].satfinite.f16x2 -> swizzled smem staging (double-buffered, handed over
//               through mbarriers);
//   warp 18     TMA stores of the staged tile (one box of 128 rows per 64 columns) at a channel offset of the
//               destination buffer (concat for free); rows past M are clipped by the TMA;
//   warps 10-17 k x k / strided / padded layers: A-gather producers - 16-byte cp.async (zero-fill for padding and
//               the K tail) straight into the 128-byte-swizzled layout the UMMA descriptor expects, each thread's
//               cp.async.mbarrier.arrive.noinc signalling the stage when its copies land.
//               1x1 layers: warps 10-12 are extra TMA issuers.  Measured on B200 (tools/umma_probe.cu part 3/4): the
//               chain try_wait -> arrive.expect_tx -> cp.async.bulk.tensor costs one thread ~600-800 cycles per
//               stage whatever the box size, while issuers in different warps scale linearly; stages are dealt
//               round-robin to n_issuers threads, n_issuers divides the ring depth so a slot always belongs to the
//               same thread (parity waits cannot alias).
// The smem ring runs across tile boundaries, so a CTA never drains its pipeline between tiles.
// With programmatic dependent launch the prologue (barriers, TMEM, constants, descriptor prefetch) of
// layer i+1 overlaps the tail of layer i; griddepcontrol.wait guards the first activation access.
#pragma once

#include "fire_common.cuh"

namespace fire {

constexpr int CONV_BM = 128;
constexpr int CONV_EPI_WARPS = 8;
constexpr int CONV_FIRST_EPI_WARP = 2;
constexpr int CONV_FIRST_HELPER_WARP = CONV_FIRST_EPI_WARP + CONV_EPI_WARPS;      // 10
constexpr int CONV_HELPER_WARPS = 8;
constexpr int CONV_HELPER_THREADS = CONV_HELPER_WARPS * 32;                       // 256
constexpr int CONV_STORE_WARP = CONV_FIRST_HELPER_WARP + CONV_HELPER_WARPS;      // 18: issues the TMA stores
constexpr int CONV_THREADS = 32 * (CONV_STORE_WARP + 1);                          // 608
constexpr int CONV_MAX_ISSUERS = 4;                                               // warp 0 + helper warps 10..12
constexpr int CONV_ROWS_PER_GATHER_THREAD = CONV_BM / (CONV_HELPER_THREADS / 8);  // 4
constexpr int CONV_A_STAGE_BYTES = CONV_BM * 128;
constexpr int CONV_MAX_COUT = 1792;
constexpr int CONV_PASS_COLS = 128;        // columns staged per epilogue pass

constexpr int CF_RELU = 1, CF_RESIDUAL = 2, CF_OUT_F32 = 4;
constexpr int CF_DBG_PHASES = 1 << 19;     // with a trace buffer: per-role cycle accounting (strip kernel)
// Work-skipping switches for timing experiments exist only in a -DFIRE_B200_SKIP_EXPERIMENTS build; in the shipped library the
// masks are 0, the tests below fold to constants and no code path can drop a gather, a store or an MMA.
#ifdef FIRE_B200_SKIP_EXPERIMENTS
constexpr int CF_DBG_NOGATHER = 1 << 16, CF_DBG_NOSTORE = 1 << 17, CF_DBG_NOMMA = 1 << 18;
#else
constexpr int CF_DBG_NOGATHER = 0, CF_DBG_NOSTORE = 0, CF_DBG_NOMMA = 0;
#endif

struct FastDiv {            // q = x / d for 0 <= x < 2^31  (mul = ceil(2^sh / d), sh = 31 + ceil(log2 d))
  uint32_t mul, sh;
};
__device__ __forceinline__ int fdiv(int x, FastDiv f) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<uint32_t>(x)) * f.mul) >> f.sh);
}

// Shared-memory carve-up, computed identically on the host (sizing) and the device.
struct ConvSmem {
  uint32_t a, b, bias, ones, zero, ident, out, bars, total;
};
__host__ __device__ inline ConvSmem conv_smem_layout(int stages, int bn, int cout, int n_res, bool pair = false) {
  ConvSmem L;
  uint32_t o = 0;
  L.a = o;     o += static_cast<uint32_t>(stages) * CONV_A_STAGE_BYTES;
  L.b = o;     o += static_cast<uint32_t>(stages) * (pair ? bn / 2 : bn) * 128;     // a CTA of a pair holds half of the weight rows
  L.bias = o;  o += static_cast<uint32_t>(cout) * 16;       // [cout] x {hi, lo, 0 x 6} fp16: K-chunk 0 of the bias operand
  L.ones = o;  o += CONV_BM * 16;                           // [128] x {1, 1, 0 x 6}: K-chunk 0 of the ones operand
  L.zero = o;  o += 256 * 16;                               // K-chunk 1 of both (all zero)
  o = (o + 1023) & ~1023u;
  L.ident = o; o += n_res ? 64 * 128 : 0;                   // 64 x 64 identity, 128-byte swizzle
  L.out = o;   o += 2 * 4 * 32 * (bn < CONV_PASS_COLS ? bn : CONV_PASS_COLS) * 2;   // [2 buffers][4 quarters] staging for the TMA store
  L.bars = o;  o += 320;
  L.total = o;
  return L;
}

struct ConvParams {
  const __half* in;  int in_ld, in_coff;      // gather mode input
  void* out;         int out_ld, out_coff;    // CF_OUT_F32 only (fp16 outputs go through tmap_out)
  const uint4* bias16;                        // [cout] x {hi, lo, 0...} fp16
  int H, W, Ho, Wo, kh, kw, stride, pad_h, pad_w;
  int cin, cout, k_real, nkb, flags, bn_tile, M_total, stages, tma_a, tmem_cols;   // tma_a: 0 = cp.async gather, 1 = tiled 2-D TMA (1x1), 2 = im2col TMA
  int m_tiles, n_tiles, pdl;
  int n_res;                    // residual K-blocks per tile (bn_tile / 64 when CF_RESIDUAL, else 0)
  int box_cols;                 // columns per TMA-store box: 64 / 32 / 16 (128B / 64B / 32B swizzle)
  long long* trace;             // debug timeline: [gridDim.x][8] globaltimer stamps (nullptr = off)
  int n_issuers;                // TMA issuing threads (1, 2 or 4; divides `stages`), stage g is issued by thread g % n_issuers
  FastDiv d_howo, d_wo, d_cin, d_kw, d_ntiles;
  int cpb;                      // tma_a == 2: 64-channel slices per tap (cin / 64)
  FastDiv d_cpb;
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CONV_TRACE(slot) do { if (p.trace) p.trace[blockIdx.x * 8 + (slot)] = globaltimer_ns(); } while (0)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// two fp32 accumulators -> packed fp16 pair (lo = a, hi = b), saturating at +-65504, optional ReLU: ONE F2FP each
__device__ __forceinline__ uint32_t cvt_pack_relu(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
  return r;
}
__device__ __forceinline__ uint32_t cvt_pack(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
  return r;
}

// one 16-column chunk (16 fp32 accumulators of this lane's row) -> 32 bytes of fp16 in the swizzled staging box
__device__ __forceinline__ void conv_stage_chunk(const uint32_t (&r)[16], bool relu, uint32_t row_addr, uint32_t u0, uint32_t swz) {
  uint4 lo, hi;
  if (relu) {
    lo = make_uint4(cvt_pack_relu(r[0], r[1]), cvt_pack_relu(r[2], r[3]), cvt_pack_relu(r[4], r[5]), cvt_pack_relu(r[6], r[7]));
    hi = make_uint4(cvt_pack_relu(r[8], r[9]), cvt_pack_relu(r[10], r[11]), cvt_pack_relu(r[12], r[13]), cvt_pack_relu(r[14], r[15]));
  } else {
    lo = make_uint4(cvt_pack(r[0], r[1]), cvt_pack(r[2], r[3]), cvt_pack(r[4], r[5]), cvt_pack(r[6], r[7]));
    hi = make_uint4(cvt_pack(r[8], r[9]), cvt_pack(r[10], r[11]), cvt_pack(r[12], r[13]), cvt_pack(r[14], r[15]));
  }
  sts128(row_addr + ((u0 ^ swz) << 4), lo);
  sts128(row_addr + (((u0 + 1) ^ swz) << 4), hi);
}

// kPair: CTA-pair instantiation (clusters of two, tcgen05 cta_group::2; TMA-fed layers without residual): the pair computes
// M = 256 x bn per MMA, each CTA loads its own A tile and HALF of the weight rows, so per MMA every SM fills and reads 16 KB
// of shared memory instead of 24 KB - the 1-CTA ceiling of DESIGN 5.7.  Only the leader (rank 0) issues MMAs; completions
// are multicast to both CTAs; the peer reports operand arrival and accumulator release with remote mbarrier arrives.
template <bool kPair>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel_t(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr bool pair = kPair;
  const uint32_t rank = pair ? cluster_ctarank() : 0u;
  const int t_first = pair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int t_step = pair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int b_rows = pair ? p.bn_tile / 2 : p.bn_tile;
  const ConvSmem L = conv_smem_layout(p.stages, p.bn_tile, p.cout, p.n_res, pair);
  const int b_stage_bytes = b_rows * 128;
  uint8_t* sA = smem + L.a;
  uint8_t* sB = smem + L.b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;      // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* bias_ready = acc_empty + 2;       // the bias operand has landed in shared memory (filled after the setup barrier)
  uint64_t* out_full = bias_ready + 1;        // [2] staging buffer written by the epilogue warps
  uint64_t* out_empty = out_full + 2;         // [2] staging buffer read by the TMA store
  uint64_t* peer_full = out_empty + 2;        // [stages] pair mode, leader: the peer CTA's operands of this stage have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + (pair ? p.stages : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = (pair ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;     // scheduling units: (pair of) M tile(s) x N tile
  // M tile of this CTA inside unit `tile`
  auto unit_mt = [&](int tile) { const int q = fdiv(tile, p.d_ntiles); return pair ? 2 * q + static_cast<int>(rank) : q; };
  const int nk_total = p.nkb + p.n_res;
  const bool out_f32 = p.flags & CF_OUT_F32;
  if (threadIdx.x == 0) CONV_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    if (p.tma_a) tma_prefetch_desc(&tmap_a);
    if (p.n_res) tma_prefetch_desc(&tmap_res);
    if (!out_f32) tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t full_count = p.tma_a ? 1u : 1u + CONV_HELPER_THREADS;
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
      // one arrival per epilogue WARP (256 per-thread arrivals on one barrier word serialise: ~250 cycles per barrier)
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], pair ? 2 * CONV_EPI_WARPS : CONV_EPI_WARPS); }
      if (pair) for (int s = 0; s < p.stages; ++s) mbar_init(&peer_full[s], 1);
      mbar_init(bias_ready, CONV_EPI_WARPS);
      for (int b = 0; b < 2; ++b) { mbar_init(&out_full[b], CONV_EPI_WARPS); mbar_init(&out_empty[b], 1); }
      fence_barrier_init();
    }
    __syncwarp();
    if (!pair) tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  if (warp >= CONV_FIRST_EPI_WARP && warp < CONV_STORE_WARP) {
    // constant MMA operands (weights-side data: safe to read before the dependency wait)
    const int t = threadIdx.x - CONV_FIRST_EPI_WARP * 32;                 // 0..511
    if (t < CONV_BM) reinterpret_cast<uint4*>(smem + L.ones)[t] = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // {1.0h, 1.0h, 0...}
    if (t < 256) reinterpret_cast<uint4*>(smem + L.zero)[t] = make_uint4(0u, 0u, 0u, 0u);
    if (p.n_res) {                                                       // 64 x 64 identity, rows of 128 bytes, SW128
      const int row = t >> 3, u = t & 7;                                 // 512 threads = 64 rows x 8 units
      const int e = row - u * 8;                                         // element index of the 1.0 inside this unit
      const uint32_t one = (e & 1) ? 0x3C000000u : 0x00003C00u;
      const int wi = (e >= 0 && e < 8) ? (e >> 1) : -1;
      *reinterpret_cast<uint4*>(smem + L.ident + sw128_offset(row, u)) =
          make_uint4(wi == 0 ? one : 0u, wi == 1 ? one : 0u, wi == 2 ? one : 0u, wi == 3 ? one : 0u);
    }
    fence_proxy_async_smem();
  }
  if (pair) {                                       // barriers of both CTAs exist before anything remote touches them
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_alloc2_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CONV_TRACE(1);
  if (p.pdl) pdl_launch_dependents();               // let the next layer start its own prologue

  const int issuer = warp == 0 ? 0
                   : (p.tma_a && warp >= CONV_FIRST_HELPER_WARP && warp < CONV_FIRST_HELPER_WARP + CONV_MAX_ISSUERS - 1)
                       ? warp - CONV_FIRST_HELPER_WARP + 1 : -1;
  if (issuer >= 0) {
    // ---------------------------------------------------------------- TMA producers (stage g belongs to issuer g % n_issuers)
    // Whole-warp loops with one ELECTED lane issuing keep addresses and descriptors in uniform registers; a
    // divergent `if (lane == 0)` body makes ptxas wrap every UTMALDG / UTCHMMA in an ELECT + R2UR + BRA.U.ANY loop.
    if (issuer < p.n_issuers) {
      const uint32_t tx = static_cast<uint32_t>(b_stage_bytes) + (p.tma_a ? CONV_A_STAGE_BYTES : 0);
      // Pass 0 (1x1 layers under PDL only): the WEIGHT halves of the first ring fill do not depend on the previous
      // layer, so they are in flight before griddepcontrol.wait; pass 1 issues everything else.
      const int prefill = (p.pdl && p.tma_a) ? p.stages : 0;
      for (int pass = prefill ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1 && p.pdl && p.tma_a) pdl_wait();          // activations / residual come from the previous layer
        int s = 0, turn = 0, g = 0;
        uint32_t ph = 0;
        for (int tile = t_first; tile < total_tiles; tile += t_step) {
          const int n0 = (tile - fdiv(tile, p.d_ntiles) * p.n_tiles) * p.bn_tile, m0 = unit_mt(tile) * CONV_BM;
          int bw = 0, bh = 0, bn_img = 0;                          // im2col: input-space base pixel of the tile's first output pixel
          if (p.tma_a == 2) {
            bn_img = fdiv(m0, p.d_howo);
            const int rem = m0 - bn_img * (p.Ho * p.Wo), ho = fdiv(rem, p.d_wo);
            bw = (rem - ho * p.Wo) * p.stride - p.pad_w;
            bh = ho * p.stride - p.pad_h;
          }
          for (int kb = 0; kb < nk_total; ++kb, ++g) {
            if (pass == 0 && g >= prefill) break;
            if (turn == issuer) {
              const bool pre = g < prefill;                      // this slot's expect_tx (and weight load) happened in pass 0
              if (!pre || pass == 0) mbar_wait(&empty[s], ph ^ 1, 11);
              if (elect_one()) {
                if (kb < p.nkb) {
                  if (!pre || pass == 0) {
                    mbar_arrive_expect_tx(&full[s], tx);
                    tma_load_2d_hint(sB + static_cast<size_t>(s) * b_stage_bytes, &tmap_w, &full[s], kb * 64, n0 + static_cast<int>(rank) * b_rows, kEvictLast);
                  }
                  if (p.tma_a == 1 && pass == 1) tma_load_2d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_a, &full[s], kb * 64, m0);
                  if (p.tma_a == 2 && pass == 1) {               // k x k layer, cin % 64 == 0: K-block kb = (tap, 64-channel slice)
                    const int tap = fdiv(kb, p.d_cpb), c0 = (kb - tap * p.cpb) * 64;
                    const int r = fdiv(tap, p.d_kw), sx = tap - r * p.kw;
                    tma_load_im2col_4d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_a, &full[s], c0, bw, bh, bn_img,
                                       static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
                  }
                } else if (pass == 1) {                           // residual K-block: [128 rows x 64 channels] of x
                  mbar_arrive_expect_tx(&full[s], CONV_A_STAGE_BYTES);
                  tma_load_2d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_res, &full[s], n0 + (kb - p.nkb) * 64, m0);
                }
              }
              __syncwarp();
              if (pass == 1 && issuer == 0 && lane == 0 && g == 0) CONV_TRACE(2);
            }
            if (++turn == p.n_issuers) turn = 0;
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
          if (pass == 0 && g >= prefill) break;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp loops, one elected lane issues)
    if (rank != 0) {
      // peer CTA of a pair: this warp issues nothing (the leader's MMAs cover both CTAs); it is the RELAY that tells the leader
      // when a stage of THIS CTA's operands has landed (relaxed arrive: the data was written by the TMA / cp.async path and
      // is read by the tensor cores; a release.cluster arrive per stage made this loop the bottleneck of the whole layer)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t_first; tile < total_tiles; tile += t_step) {
        for (int kb = 0; kb < nk_total; ++kb) {
          mbar_wait(&full[s], ph, 21);
          if (lane == 0) mbar_arrive_remote_relaxed(&peer_full[s], 0u);
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    } else {                                           // leader (or single CTA)
      const uint32_t idesc = umma_idesc_f16(pair ? 2 * CONV_BM : CONV_BM, p.bn_tile);
      const uint32_t idesc64 = umma_idesc_f16(CONV_BM, 64);
      const uint32_t ones_addr = smem_u32(smem + L.ones), zero_addr = smem_u32(smem + L.zero);
      const uint32_t bias_addr = smem_u32(smem + L.bias), ident_addr = smem_u32(smem + L.ident);
      const uint64_t ones_desc = umma_desc_nosw(ones_addr, zero_addr - ones_addr, 128);
      const bool do_mma = !(p.flags & CF_DBG_NOMMA);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      int lt = 0, s = 0;
      uint32_t ph = 0;
      long long tw[3] = {0, 0, 0};                               // cycles: waiting for a free accumulator / for operands / issuing
      const bool prof = p.trace != nullptr && (p.flags & CF_DBG_PHASES);
      for (int tile = t_first; tile < total_tiles; tile += t_step, ++lt) {
        const int buf = lt & 1;
        const int n0 = (tile - fdiv(tile, p.d_ntiles) * p.n_tiles) * p.bn_tile;
        long long c0 = prof ? clock64() : 0, c1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1, 15);
        tc_fence_after();
        if (prof) { c1 = clock64(); tw[0] += c1 - c0; c0 = c1; }
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * p.bn_tile);
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (prof) c0 = clock64();
          mbar_wait(&full[s], ph, 12);
          if (pair) mbar_wait(&peer_full[s], ph, 20);
          tc_fence_after();
          if (prof) { c1 = clock64(); tw[1] += c1 - c0; c0 = c1; }
          if (lt == 0 && kb == 0 && lane == 0) CONV_TRACE(3);
          if (elect_one()) {
            const uint32_t a0 = a_base + static_cast<uint32_t>(s * CONV_A_STAGE_BYTES);
            const uint32_t b0 = b_base + static_cast<uint32_t>(s * b_stage_bytes);
            if (do_mma) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (pair) umma_f16_2cta(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                else umma_f16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            if (pair) umma_commit_2cta(&empty[s]); else umma_commit(&empty[s]);
          }
          __syncwarp();
          if (prof) { c1 = clock64(); tw[2] += c1 - c0; }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        for (int j = 0; j < p.n_res; ++j) {                     // D[:, 64j .. 64j+63] += R_j * I
          if (prof) c0 = clock64();
          mbar_wait(&full[s], ph, 18);
          tc_fence_after();
          if (prof) { c1 = clock64(); tw[1] += c1 - c0; }
          if (elect_one()) {
            const uint32_t a0 = a_base + static_cast<uint32_t>(s * CONV_A_STAGE_BYTES);
            if (do_mma) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(d + static_cast<uint32_t>(j * 64), umma_desc_sw128(a0 + k * 32), umma_desc_sw128(ident_addr + k * 32), idesc64, 1u);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (lt == 0) mbar_wait(bias_ready, 0, 19);
        if (elect_one()) {                                      // D += ones * bias^T, then publish the accumulator
          const uint32_t b_addr = bias_addr + static_cast<uint32_t>(n0) * 16;     // pair: each CTA keeps ITS half of the tile's rows here
          if (pair) {
            umma_f16_2cta(d, ones_desc, umma_desc_nosw(b_addr, zero_addr - b_addr, 128), idesc, 1u);
            umma_commit_2cta(&acc_full[buf]);
          } else {
            umma_f16(d, ones_desc, umma_desc_nosw(b_addr, zero_addr - b_addr, 128), idesc, 1u);
            umma_commit(&acc_full[buf]);
          }
        }
        __syncwarp();
      }
      if (lane == 0) CONV_TRACE(4);
      if (prof && lane == 0) {
        long long* q = p.trace + 8 * 148 + blockIdx.x * 8;
        q[0] = tw[0]; q[1] = tw[1]; q[2] = tw[2]; q[7] = lt;
      }
    }
  } else if (warp < CONV_FIRST_HELPER_WARP) {
    // ---------------------------------------------------------------- epilogue (8 warps, 2 per TMEM lane quarter)
    const int quarter = warp & 3, half = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const bool relu = p.flags & CF_RELU;
    const int pass_cols = min(p.bn_tile, CONV_PASS_COLS);
    const int rowbytes = p.box_cols * 2, box_bytes = CONV_BM * rowbytes, chunks_per_box = p.box_cols >> 4;
    const int m_row = quarter * 32 + lane;                      // accumulator row of this thread
    const uint32_t swz = p.box_cols == 64 ? (m_row & 7) : p.box_cols == 32 ? ((m_row >> 1) & 3) : ((m_row >> 2) & 1);
    const uint32_t stage_buf_bytes = static_cast<uint32_t>(CONV_BM * pass_cols * 2);                // one staging buffer: [boxes][128 rows][rowbytes]
    const uint32_t stage0 = smem_u32(smem + L.out);
    int pc = 0;                                                 // pass counter: staging buffer pc & 1
    const bool do_store = !(p.flags & CF_DBG_NOSTORE);
    {
      // bias operand (a weight: no dependency wait needed); only needed by the LAST MMA of the first tile, so it is
      // fetched here, off the critical path of the prologue
      uint4* s_bias = reinterpret_cast<uint4*>(smem + L.bias);
      for (int i = threadIdx.x - CONV_FIRST_EPI_WARP * 32; i < p.cout; i += CONV_EPI_WARPS * 32) {
        // pair: the MMA descriptor addresses row j of an N tile at the same offset in both CTAs; the peer keeps rows bn/2.. there
        const int j = i % p.bn_tile, src = pair ? i - j + (j + static_cast<int>(rank) * b_rows) % p.bn_tile : i;
        s_bias[i] = __ldg(p.bias16 + src);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bias_ready);
    }
    if (p.pdl && out_f32) pdl_wait();                           // output writes must not overtake readers of the previous layers
    int lt = 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step, ++lt) {
      const int buf = lt & 1;
      const int n0 = (tile - fdiv(tile, p.d_ntiles) * p.n_tiles) * p.bn_tile, m0 = unit_mt(tile) * CONV_BM;
      mbar_wait(&acc_full[buf], (lt >> 1) & 1, 14);
      tc_fence_after();
      if (lt == 0 && threadIdx.x == CONV_FIRST_EPI_WARP * 32) CONV_TRACE(5);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.bn_tile);
      if (out_f32) {
        // fp32 output (the bottleneck GEMM only): thread = row, direct 16-byte stores
        const int m = m0 + quarter * 32 + lane;
        const int n_chunks = p.bn_tile >> 4;
        for (int c = half; c < n_chunks; c += 2) {
          uint32_t r[16];
          __syncwarp();
          tmem_ld_32x16(taddr + static_cast<uint32_t>(c * 16), r);
          tmem_ld_wait(r);
          if (m < p.M_total && do_store) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_ld + p.out_coff + n0 + c * 16);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float4 v = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
              if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              op[e] = v;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (rank != 0) mbar_arrive_remote_relaxed(&acc_empty[buf], 0u); else mbar_arrive(&acc_empty[buf]); }
        continue;
      }
      for (int cg = 0; cg < p.bn_tile; cg += CONV_PASS_COLS, ++pc) {
        const int cols = min(CONV_PASS_COLS, p.bn_tile - cg);
        const int n_chunks = cols >> 4;
        const int ob = pc & 1;
        const uint32_t my_row = stage0 + static_cast<uint32_t>(ob) * stage_buf_bytes + static_cast<uint32_t>(m_row * rowbytes);
        mbar_wait(&out_empty[ob], ((pc >> 1) & 1) ^ 1, 16);     // the store issued two passes ago has left this staging buffer
        uint32_t ra[16], rb[16];
        int c = half;
        __syncwarp();
        if (c < n_chunks) tmem_ld_32x16(taddr + static_cast<uint32_t>(cg + c * 16), ra);
        for (; c < n_chunks; c += 4) {
          // chunk c (in ra), prefetch c + 2 (into rb)
          tmem_ld_wait(ra);
          __syncwarp();
          if (c + 2 < n_chunks) tmem_ld_32x16(taddr + static_cast<uint32_t>(cg + (c + 2) * 16), rb);
          {
            const int box = c / chunks_per_box, cb = c - box * chunks_per_box;
            conv_stage_chunk(ra, relu, my_row + static_cast<uint32_t>(box * box_bytes), static_cast<uint32_t>(2 * cb), swz);
          }
          if (c + 2 < n_chunks) {
            tmem_ld_wait(rb);
            __syncwarp();
            if (c + 4 < n_chunks) tmem_ld_32x16(taddr + static_cast<uint32_t>(cg + (c + 4) * 16), ra);
            const int c2 = c + 2;
            const int box = c2 / chunks_per_box, cb = c2 - box * chunks_per_box;
            conv_stage_chunk(rb, relu, my_row + static_cast<uint32_t>(box * box_bytes), static_cast<uint32_t>(2 * cb), swz);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();                               // staging writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (cg + CONV_PASS_COLS >= p.bn_tile) {             // accumulator fully read: hand it back (to the leader's MMA warp)
            if (rank != 0) mbar_arrive_remote_relaxed(&acc_empty[buf], 0u); else mbar_arrive(&acc_empty[buf]);
          }
          mbar_arrive(&out_full[ob]);
        }
      }
    }
    if (threadIdx.x == CONV_FIRST_EPI_WARP * 32) CONV_TRACE(6);
  } else if (warp == CONV_STORE_WARP) {
    // ---------------------------------------------------------------- TMA store warp
    if (!out_f32) {
      const int pass_cols = min(p.bn_tile, CONV_PASS_COLS);
      const int box_bytes = CONV_BM * p.box_cols * 2;
      const uint32_t stage_buf_bytes = static_cast<uint32_t>(CONV_BM * pass_cols * 2);
      const uint32_t stage0 = smem_u32(smem + L.out);
      const bool do_store = !(p.flags & CF_DBG_NOSTORE);
      if (p.pdl) pdl_wait();                                    // output writes must not overtake readers of the previous layers
      int pc = 0;
      for (int tile = t_first; tile < total_tiles; tile += t_step) {
        const int n0 = (tile - fdiv(tile, p.d_ntiles) * p.n_tiles) * p.bn_tile, m0 = unit_mt(tile) * CONV_BM;
        for (int cg = 0; cg < p.bn_tile; cg += CONV_PASS_COLS, ++pc) {
          const int cols = min(CONV_PASS_COLS, p.bn_tile - cg);
          const int ob = pc & 1;
          mbar_wait(&out_full[ob], (pc >> 1) & 1, 17);
          if (elect_one()) {
            if (do_store) {
              const int n_boxes = cols / p.box_cols;
              for (int b = 0; b < n_boxes; ++b)
                tma_store_2d(&tmap_out, stage0 + static_cast<uint32_t>(ob) * stage_buf_bytes + static_cast<uint32_t>(b * box_bytes),
                             n0 + cg + b * p.box_cols, m0);
            }
            bulk_commit_group();
            bulk_wait_read_all();                               // staging buffer read: hand it back
            mbar_arrive(&out_empty[ob]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) bulk_wait_all();                         // stores complete before the CTA exits
      __syncwarp();
    }
  } else if (!p.tma_a && warp < CONV_STORE_WARP) {
    // ---------------------------------------------------------------- A gather producers (8 warps)
    const int g = threadIdx.x - CONV_FIRST_HELPER_WARP * 32;   // 0..255
    const int chunk = g & 7, rbase = g >> 3;                    // 8 lanes cover one 128-byte row; rows rbase + 32*i
    const uint32_t sw_const = smem_u32(sA) + static_cast<uint32_t>(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));   // (row & 7) == (rbase & 7)
    const int HoWo = p.Ho * p.Wo;
    const bool do_copy = !(p.flags & CF_DBG_NOGATHER);
    const __half* __restrict__ inp = p.in;
    if (p.pdl) pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step) {
      const int m0 = unit_mt(tile) * CONV_BM;
      int base_off[CONV_ROWS_PER_GATHER_THREAD];
      uint32_t vmask[CONV_ROWS_PER_GATHER_THREAD];  // bits 0..7: tap rows r with ih in range; bits 8..15: tap cols s with iw in range
#pragma unroll
      for (int i = 0; i < CONV_ROWS_PER_GATHER_THREAD; ++i) {
        const int m = m0 + rbase + 32 * i;
        const int n = fdiv(m, p.d_howo), rem = m - n * HoWo;
        const int ho = fdiv(rem, p.d_wo), wo = rem - ho * p.Wo;
        const int ih0 = ho * p.stride - p.pad_h, iw0 = wo * p.stride - p.pad_w;
        base_off[i] = ((n * p.H + ih0) * p.W + iw0) * p.in_ld + p.in_coff;
        // taps r in [max(0,-ih0), min(kh, H-ih0)) and s in [max(0,-iw0), min(kw, W-iw0)) read inside the image
        const int r_lo = max(0, -ih0), r_hi = min(p.kh, p.H - ih0), c_lo = max(0, -iw0), c_hi = min(p.kw, p.W - iw0);
        const uint32_t rm = r_hi > r_lo ? ((1u << r_hi) - (1u << r_lo)) : 0u;
        const uint32_t cm = c_hi > c_lo ? ((1u << c_hi) - (1u << c_lo)) : 0u;
        vmask[i] = m < p.M_total ? (rm | (cm << 8)) : 0u;
      }
      int k = chunk * 8;
      for (int kb = 0; kb < p.nkb; ++kb, k += 64) {
        const int tap = fdiv(k, p.d_cin), c = k - tap * p.cin;
        const int r = fdiv(tap, p.d_kw), sx = tap - r * p.kw;
        const uint32_t need = k < p.k_real ? ((1u << r) | (256u << sx)) : 0xFFFFFFFFu;   // K tail: never valid
        const int tap_off = (r * p.W + sx) * p.in_ld + c;
        const uint32_t a_s = sw_const + static_cast<uint32_t>(s * CONV_A_STAGE_BYTES);
        mbar_wait(&empty[s], ph ^ 1, 13);
        if (do_copy) {
#pragma unroll
          for (int i = 0; i < CONV_ROWS_PER_GATHER_THREAD; ++i) {
            const bool ok = (vmask[i] & need) == need;
            cp_async_16(a_s + i * 32 * 128, inp + (ok ? base_off[i] + tap_off : 0), ok);
          }
        }
        cp_async_mbar_arrive_noinc(&full[s]);         // counted arrival fires when this thread's copies have landed
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    cp_async_wait_all();                              // do not exit with copies in flight
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc2_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols)); else tmem_dealloc_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  if (threadIdx.x == 0) CONV_TRACE(7);
}

}  // namespace fire
