// conv_igemm.cuh - persistent implicit-GEMM convolution on tcgen05 (K2 of DESIGN.md).
//
// One launch = one Conv(+folded BN)(+bias)(+residual)(+ReLU) layer of the FaceNet plan
// (reference: the onnxruntime session at facenet_gpu.py:127; graph in SURVEY App. A).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]      m = output pixel (b, ho, wo), n = output channel,
//                                          k = (tap r,s ; input channel c), tap-major
//
// Persistent CTAs (one per SM) walk the (m-tile, n-tile) list; the warp roles run concurrently:
//   warp 0      TMA producer: weight K-blocks (always) and, for 1x1/stride-1 layers, the activation
//               K-blocks too (the activation matrix [M, C] is then a plain 2-D tensor);
//   warps 14-16 extra TMA producers.  Measured on B200 (tools/umma_probe.cu part 3/4): the chain
//               try_wait -> arrive.expect_tx -> cp.async.bulk.tensor costs one thread ~600-800 cycles per
//               stage whatever the box size, i.e. ONE issuing thread tops out near 20 B/clk/SM, while
//               issuers in different warps scale linearly (4 x 16 KB boxes reach the 17 TB/s L2 limit).
//               Stages are therefore dealt round-robin to n_issuers threads in different warps; n_issuers
//               divides the ring depth, so a slot always belongs to the same thread (parity waits cannot alias);
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=bn_tile, K=16) into one of TWO
//               TMEM accumulators, so the epilogue of tile t overlaps the main loop of tile t+1;
//   warps 2-5   epilogue: tcgen05.ld -> +bias (smem) -> +residual (smem) -> ReLU -> fp16 -> per-warp
//               smem staging -> row-contiguous 16-byte global stores at a channel offset of the
//               destination buffer (concat for free).  Thread = accumulator row only while talking
//               to TMEM; every global access is coalesced along the channel dimension;
//   warps 6-13  helpers.  k x k / strided / padded layers: A-gather producers - 16-byte cp.async
//               (zero-fill for padding and the K tail) straight into the 128-byte-swizzled layout the
//               UMMA descriptor expects, each thread's cp.async.mbarrier.arrive.noinc signalling the
//               stage when its copies land; index math uses precomputed magic-number division and
//               per-row tap-validity masks.  1x1 layers with a residual: the same warps prefetch the
//               residual tile of the NEXT tile into a double-buffered smem stage with coalesced cp.async.
// The smem ring runs across tile boundaries, so a CTA never drains its pipeline between tiles.
// With programmatic dependent launch the prologue (barriers, TMEM, bias, descriptor prefetch) of
// layer i+1 overlaps the tail of layer i; griddepcontrol.wait guards the first activation access.
#pragma once

#include "fire_common.cuh"

namespace fire {

constexpr int CONV_BM = 128;
constexpr int CONV_HELPER_WARPS = 8;
constexpr int CONV_HELPER_THREADS = CONV_HELPER_WARPS * 32;            // 256
constexpr int CONV_EXTRA_ISSUERS = 3;                                   // warps 14..16: additional TMA issuing threads
constexpr int CONV_THREADS = 64 + 128 + CONV_HELPER_THREADS + 32 * CONV_EXTRA_ISSUERS;   // 544
constexpr int CONV_ROWS_PER_GATHER_THREAD = CONV_BM / (CONV_HELPER_THREADS / 8);   // 4
constexpr int CONV_A_STAGE_BYTES = CONV_BM * 128;
constexpr int CONV_MAX_COUT = 1792;
constexpr int CONV_STAGE_COLS = 128;       // columns staged per epilogue pass (sOut width)

constexpr int CF_RELU = 1, CF_RESIDUAL = 2, CF_OUT_F32 = 4;
constexpr int CF_DBG_NOGATHER = 1 << 16, CF_DBG_NOSTORE = 1 << 17, CF_DBG_NOMMA = 1 << 18;   // timing experiments only (wrong results)

struct FastDiv {            // q = x / d for 0 <= x < 2^31  (mul = ceil(2^sh / d), sh = 31 + ceil(log2 d))
  uint32_t mul, sh;
};
__device__ __forceinline__ int fdiv(int x, FastDiv f) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<uint32_t>(x)) * f.mul) >> f.sh);
}

struct ConvParams {
  const __half* in;  int in_ld, in_coff;
  void* out;         int out_ld, out_coff;
  const __half* res; int res_ld, res_coff;
  const float* bias;
  int H, W, Ho, Wo, kh, kw, stride, pad_h, pad_w;
  int cin, cout, k_real, nkb, flags, bn_tile, M_total, stages, tma_a, tmem_cols;
  int m_tiles, n_tiles, pdl;
  int res_smem;                 // residual tile prefetched into smem by the helper warps (tma_a && residual)
  long long* trace;             // debug timeline: [gridDim.x][8] globaltimer stamps (nullptr = off)
  int n_issuers;                // TMA issuing threads (1, 2 or 4; divides `stages`), stage g is issued by thread g % n_issuers
  FastDiv d_howo, d_wo, d_cin, d_kw, d_unit_res, d_unit_out, d_ntiles;
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

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// one 16-column chunk of one accumulator row: + bias (4 x LDS.128), + residual -> 16 floats in v[]
__device__ __forceinline__ void conv_chunk_math(const uint32_t (&r)[16], const uint4 (&q)[2], uint32_t s_bias_addr, bool has_res,
                                                float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 b = lds128(s_bias_addr + j * 16);
    v[4 * j] = __uint_as_float(r[4 * j]) + __uint_as_float(b.x);
    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + __uint_as_float(b.y);
    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + __uint_as_float(b.z);
    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + __uint_as_float(b.w);
  }
  if (has_res) {
    const __half2* h0 = reinterpret_cast<const __half2*>(&q[0]);
    const __half2* h1 = reinterpret_cast<const __half2*>(&q[1]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
      v[2 * j] += f0.x; v[2 * j + 1] += f0.y;
      v[8 + 2 * j] += f1.x; v[8 + 2 * j + 1] += f1.y;
    }
  }
}
// fp32 pair -> packed fp16 with ReLU (lo = 0) or without (lo = -65504) and saturation at +65504, done on the packed pair
__device__ __forceinline__ uint32_t pack_f16x2_clamp(float a, float b, __half2 lo, __half2 hi) {
  __half2 h = __hmin2(__hmax2(__floats2half2_rn(a, b), lo), hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.bn_tile * 128;
  const int sw_cols = min(p.bn_tile, CONV_STAGE_COLS);             // columns per staging pass
  const int out_pitch = sw_cols * 2 + 16;                          // bytes; (pitch/16) odd -> conflict-free row-per-lane access
  const int res_pitch = p.bn_tile * 2 + 16;
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.stages) * CONV_A_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(sB + static_cast<size_t>(p.stages) * b_stage_bytes);
  uint8_t* sOut = reinterpret_cast<uint8_t*>(s_bias + CONV_MAX_COUT);                  // [4 warps][32 rows][out_pitch]
  uint8_t* sRes = sOut + 4 * 32 * out_pitch;                                            // [2][128 rows][res_pitch] (res_smem only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRes + (p.res_smem ? 2 * CONV_BM * res_pitch : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;      // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* res_full = acc_empty + 2;         // [2]
  uint64_t* res_empty = res_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  if (threadIdx.x == 0) CONV_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    if (p.tma_a) tma_prefetch_desc(&tmap_a);
  }
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t full_count = p.tma_a ? 1u : 1u + CONV_HELPER_THREADS;
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128);
        mbar_init(&res_full[b], CONV_HELPER_THREADS); mbar_init(&res_empty[b], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  if (warp >= 2 && warp < 6) {                      // bias is a weight: safe to read before the dependency wait
    for (int i = threadIdx.x - 64; i < p.cout; i += 128) s_bias[i] = __ldg(p.bias + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CONV_TRACE(1);
  if (p.pdl) pdl_launch_dependents();               // let the next layer start its own prologue

  const int issuer = warp == 0 ? 0 : (warp >= 14 ? warp - 13 : -1);
  if (issuer >= 0) {
    // ---------------------------------------------------------------- TMA producers (stage g belongs to issuer g % n_issuers)
    if (lane == 0 && issuer < p.n_issuers) {
      const uint32_t tx = static_cast<uint32_t>(b_stage_bytes) + (p.tma_a ? CONV_A_STAGE_BYTES : 0);
      bool waited = !p.pdl || !p.tma_a;
      int s = 0, turn = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = fdiv(tile, p.d_ntiles);
        const int n0 = (tile - mt * p.n_tiles) * p.bn_tile, m0 = mt * CONV_BM;
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (turn == issuer) {
            mbar_wait(&empty[s], ph ^ 1, 11);
            mbar_arrive_expect_tx(&full[s], tx);
            tma_load_2d_hint(sB + static_cast<size_t>(s) * b_stage_bytes, &tmap_w, &full[s], kb * 64, n0, kEvictLast);
            if (p.tma_a) {
              if (!waited) { pdl_wait(); waited = true; }     // activations come from the previous layer
              tma_load_2d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_a, &full[s], kb * 64, m0);
            }
            if (issuer == 0 && tile == blockIdx.x && kb == 0) CONV_TRACE(2);
          }
          if (++turn == p.n_issuers) turn = 0;
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(CONV_BM, p.bn_tile);
      int lt = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1, 15);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * p.bn_tile);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full[s], ph, 12);
          tc_fence_after();
          if (lt == 0 && kb == 0) CONV_TRACE(3);
          const uint32_t a0 = smem_u32(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES);
          const uint32_t b0 = smem_u32(sB + static_cast<size_t>(s) * b_stage_bytes);
          if (!(p.flags & CF_DBG_NOMMA))
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
      CONV_TRACE(4);
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue (4 warps)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool relu = p.flags & CF_RELU, has_res = p.flags & CF_RESIDUAL, out_f32 = p.flags & CF_OUT_F32;
    const bool res_smem = p.res_smem != 0;
    const uint32_t s_bias_u32 = smem_u32(s_bias);
    const __half2 h_lo = __float2half2_rn(relu ? 0.f : -65504.f), h_hi = __float2half2_rn(65504.f);
    const uint32_t my_out = smem_u32(sOut) + static_cast<uint32_t>((warp - 2) * 32 * out_pitch);
    // coalesced copy-out geometry: 16-byte units, U per row; lane starts at unit `lane` and advances 32 units per step
    const int U = sw_cols >> 3;
    const int u_row0 = fdiv(lane, p.d_unit_out), u_col0 = lane - u_row0 * U;
    const int u_drow = fdiv(32, p.d_unit_out), u_dcol = 32 - u_drow * U;
    if (p.pdl) pdl_wait();                            // residual reads / output writes depend on earlier layers
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int mt = fdiv(tile, p.d_ntiles);
      const int n0 = (tile - mt * p.n_tiles) * p.bn_tile, m0 = mt * CONV_BM;
      const int m = m0 + row;
      const bool mvalid = m < p.M_total;
      const __half* resp = has_res && !res_smem && mvalid ? p.res + static_cast<size_t>(m) * p.res_ld + p.res_coff + n0 : nullptr;
      const uint32_t my_res = smem_u32(sRes) + static_cast<uint32_t>(buf * CONV_BM * res_pitch + row * res_pitch);
      if (res_smem) mbar_wait(&res_full[buf], (lt >> 1) & 1, 16);
      mbar_wait(&acc_full[buf], (lt >> 1) & 1, 14);
      tc_fence_after();
      if (lt == 0 && threadIdx.x == 64) CONV_TRACE(5);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.bn_tile);
      for (int cg = 0; cg < p.bn_tile; cg += sw_cols) {
        const int n_chunks = sw_cols >> 4;
        uint32_t ra[16], rb[16];
        __syncwarp();
        tmem_ld_32x16(taddr + static_cast<uint32_t>(cg), ra);
        for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int cc = c + j;
            if (cc >= n_chunks) break;               // warp-uniform
            if (j) tmem_ld_wait(rb); else tmem_ld_wait(ra);
            __syncwarp();
            if (cc + 1 < n_chunks) {
              if (j) tmem_ld_32x16(taddr + static_cast<uint32_t>(cg + (cc + 1) * 16), ra);
              else tmem_ld_32x16(taddr + static_cast<uint32_t>(cg + (cc + 1) * 16), rb);
            }
            const int col = cg + cc * 16;            // column inside the tile
            uint4 q[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (res_smem) { q[0] = lds128(my_res + col * 2); q[1] = lds128(my_res + col * 2 + 16); }
            else if (resp) { q[0] = __ldg(reinterpret_cast<const uint4*>(resp + col)); q[1] = __ldg(reinterpret_cast<const uint4*>(resp + col) + 1); }
            float v[16];
            conv_chunk_math(j ? rb : ra, q, s_bias_u32 + static_cast<uint32_t>((n0 + col) * 4), has_res, v);
            if (out_f32) {
              if (relu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
              }
              if (mvalid) {
                float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_ld + p.out_coff + n0 + col);
#pragma unroll
                for (int e = 0; e < 4; ++e) op[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
              }
            } else {
              const uint32_t dst = my_out + static_cast<uint32_t>(lane * out_pitch + cc * 32);
              sts128(dst, make_uint4(pack_f16x2_clamp(v[0], v[1], h_lo, h_hi), pack_f16x2_clamp(v[2], v[3], h_lo, h_hi),
                                     pack_f16x2_clamp(v[4], v[5], h_lo, h_hi), pack_f16x2_clamp(v[6], v[7], h_lo, h_hi)));
              sts128(dst + 16, make_uint4(pack_f16x2_clamp(v[8], v[9], h_lo, h_hi), pack_f16x2_clamp(v[10], v[11], h_lo, h_hi),
                                          pack_f16x2_clamp(v[12], v[13], h_lo, h_hi), pack_f16x2_clamp(v[14], v[15], h_lo, h_hi)));
            }
          }
        }
        if (cg + sw_cols >= p.bn_tile) {             // accumulator and residual stage fully consumed: hand them back early
          __syncwarp();
          tc_fence_before();
          mbar_arrive(&acc_empty[buf]);
          if (res_smem) mbar_arrive(&res_empty[buf]);
        }
        if (!out_f32) {
          __syncwarp();
          // staged sub-tile (32 rows x sw_cols) -> global, 16 bytes per lane, lanes contiguous along the row
          const int rows_left = p.M_total - (m0 + quarter * 32);
          int ur = u_row0, uc = u_col0;
          uint32_t sp = my_out + static_cast<uint32_t>(u_row0 * out_pitch + u_col0 * 16);
          uint8_t* gp = reinterpret_cast<uint8_t*>(static_cast<__half*>(p.out) + static_cast<size_t>(m0 + quarter * 32 + u_row0) * p.out_ld +
                                                   p.out_coff + n0 + cg + u_col0 * 8);
          const uint32_t s_step = static_cast<uint32_t>(u_drow * out_pitch + u_dcol * 16), s_wrap = static_cast<uint32_t>(out_pitch - U * 16);
          const long long g_step = static_cast<long long>(u_drow) * p.out_ld * 2 + u_dcol * 16, g_wrap = static_cast<long long>(p.out_ld) * 2 - U * 16;
          const bool do_store = !(p.flags & CF_DBG_NOSTORE);
          for (int it2 = 0; it2 < U; ++it2) {
            if (ur < rows_left && do_store) *reinterpret_cast<uint4*>(gp) = lds128(sp);
            ur += u_drow; uc += u_dcol; sp += s_step; gp += g_step;
            if (uc >= U) { uc -= U; ++ur; sp += s_wrap; gp += g_wrap; }
          }
          __syncwarp();
        }
      }
    }
    if (threadIdx.x == 64) CONV_TRACE(6);
  } else if (!p.tma_a) {
    // ---------------------------------------------------------------- A gather producers (8 warps)
    const int g = threadIdx.x - 192;                // 0..255
    const int chunk = g & 7, rbase = g >> 3;        // 8 lanes cover one 128-byte row; rows rbase + 32*i
    const uint32_t sw_const = smem_u32(sA) + static_cast<uint32_t>(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));   // (row & 7) == (rbase & 7)
    const int HoWo = p.Ho * p.Wo;
    const bool do_copy = !(p.flags & CF_DBG_NOGATHER);
    const __half* __restrict__ inp = p.in;
    if (p.pdl) pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = fdiv(tile, p.d_ntiles) * CONV_BM;
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
  } else if (p.res_smem) {
    // ---------------------------------------------------------------- residual prefetchers (8 warps, 1x1 layers)
    const int g = threadIdx.x - 192;
    const int U = p.bn_tile >> 3;                     // 16-byte units per residual row
    const int total_units = CONV_BM * U;
    if (p.pdl) pdl_wait();
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int mt = fdiv(tile, p.d_ntiles);
      const int n0 = (tile - mt * p.n_tiles) * p.bn_tile, m0 = mt * CONV_BM;
      mbar_wait(&res_empty[buf], ((lt >> 1) & 1) ^ 1, 17);
      const uint32_t dst0 = smem_u32(sRes) + static_cast<uint32_t>(buf * CONV_BM * res_pitch);
      for (int u = g; u < total_units; u += CONV_HELPER_THREADS) {
        const int rr = fdiv(u, p.d_unit_res), cc = u - rr * U;
        const bool ok = m0 + rr < p.M_total;
        const __half* src = ok ? p.res + static_cast<size_t>(m0 + rr) * p.res_ld + p.res_coff + n0 + cc * 8 : p.res;
        cp_async_16(dst0 + static_cast<uint32_t>(rr * res_pitch + cc * 16), src, ok);
      }
      cp_async_mbar_arrive_noinc(&res_full[buf]);
    }
    cp_async_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  if (threadIdx.x == 0) CONV_TRACE(7);
}

}  // namespace fire
