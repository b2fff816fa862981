// conv_igemm.cuh - persistent implicit-GEMM convolution on tcgen05 (K2 of DESIGN.md).
//
// One launch = one Conv(+folded BN)(+bias)(+residual)(+ReLU) layer of the FaceNet plan
// (reference: the onnxruntime session at facenet_gpu.py:127; graph in SURVEY App. A).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]      m = output pixel (b, ho, wo), n = output channel,
//                                          k = (tap r,s ; input channel c), tap-major
//
// Persistent CTAs (one per SM) walk the (m-tile, n-tile) list; five warp roles run concurrently:
//   warp 0      TMA producer: weight K-blocks (always) and, for 1x1/stride-1 layers, the activation
//               K-blocks too (the activation matrix [M, C] is then a plain 2-D tensor);
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=bn_tile, K=16) into one of TWO
//               TMEM accumulators, so the epilogue of tile t overlaps the main loop of tile t+1;
//   warps 2-5   epilogue: tcgen05.ld -> +bias (staged in smem) -> +residual -> ReLU -> fp16/fp32
//               store at a channel offset of the destination buffer (concat for free); the TMEM and
//               residual loads of chunk c+1 are in flight while chunk c is finished;
//   warps 6-9   A-gather producers for k x k / strided / padded layers: 16-byte cp.async copies
//               (zero-fill for padding and the K tail) straight into the 128-byte-swizzled layout
//               the UMMA descriptor expects; each thread's cp.async.mbarrier.arrive.noinc signals the stage
//               when its copies land, so the producers run a full ring ahead without ever waiting on data.
// The smem ring runs across tile boundaries, so a CTA never drains its pipeline between tiles.
// With programmatic dependent launch the prologue (barriers, TMEM, bias, descriptor prefetch) of
// layer i+1 overlaps the tail of layer i; griddepcontrol.wait guards the first activation access.
#pragma once

#include "fire_common.cuh"

namespace fire {

constexpr int CONV_BM = 128;
constexpr int CONV_THREADS = 320;
constexpr int CONV_A_STAGE_BYTES = CONV_BM * 128;
constexpr int CONV_RES_DEPTH = 4;          // residual chunks prefetched ahead of the chunk being finished
constexpr int CONV_MAX_COUT = 1792;

constexpr int CF_RELU = 1, CF_RESIDUAL = 2, CF_OUT_F32 = 4, CF_GATHER_L1 = 256;

struct ConvParams {
  const __half* in;  int in_ld, in_coff;
  void* out;         int out_ld, out_coff;
  const __half* res; int res_ld, res_coff;
  const float* bias;
  int H, W, Ho, Wo, kh, kw, stride, pad_h, pad_w;
  int cin, cout, k_real, nkb, flags, bn_tile, M_total, stages, tma_a, tmem_cols;
  int m_tiles, n_tiles, pdl;
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// finish one 16-column chunk of one output row: bias, residual, ReLU, convert, store
__device__ __forceinline__ void conv_finish_chunk(const ConvParams& p, const uint32_t (&r)[16], const uint4 (&q)[2],
                                                  const float* s_bias, int m, int n, bool relu, bool has_res, bool out_f32) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + s_bias[n + j];
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
  if (relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (out_f32) {
    float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_ld + p.out_coff + n);
#pragma unroll
    for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* op = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + static_cast<size_t>(m) * p.out_ld + p.out_coff + n);
    op[0] = make_uint4(pack_f16x2_sat(v[0], v[1]), pack_f16x2_sat(v[2], v[3]), pack_f16x2_sat(v[4], v[5]), pack_f16x2_sat(v[6], v[7]));
    op[1] = make_uint4(pack_f16x2_sat(v[8], v[9]), pack_f16x2_sat(v[10], v[11]), pack_f16x2_sat(v[12], v[13]), pack_f16x2_sat(v[14], v[15]));
  }
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.bn_tile * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.stages) * CONV_A_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(sB + static_cast<size_t>(p.stages) * b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + CONV_MAX_COUT);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;      // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    if (p.tma_a) tma_prefetch_desc(&tmap_a);
  }
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t full_count = p.tma_a ? 1u : 1u + 128u;
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
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
  if (p.pdl) pdl_launch_dependents();               // let the next layer start its own prologue

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(b_stage_bytes) + (p.tma_a ? CONV_A_STAGE_BYTES : 0);
      bool waited = !p.pdl || !p.tma_a;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * p.bn_tile, m0 = (tile / p.n_tiles) * CONV_BM;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1, 11);
          mbar_arrive_expect_tx(&full[s], tx);
          tma_load_2d_hint(sB + static_cast<size_t>(s) * b_stage_bytes, &tmap_w, &full[s], kb * 64, n0, kEvictLast);
          if (p.tma_a) {
            if (!waited) { pdl_wait(); waited = true; }     // activations come from the previous layer
            tma_load_2d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_a, &full[s], kb * 64, m0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(CONV_BM, p.bn_tile);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1, 15);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * p.bn_tile);
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full[s], ph, 12);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES);
          const uint32_t b0 = smem_u32(sB + static_cast<size_t>(s) * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue (4 warps)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool relu = p.flags & CF_RELU, has_res = p.flags & CF_RESIDUAL, out_f32 = p.flags & CF_OUT_F32;
    const int n_chunks = p.bn_tile >> 4;
    if (p.pdl) pdl_wait();                            // residual reads / output writes depend on earlier layers
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int n0 = (tile % p.n_tiles) * p.bn_tile, m0 = (tile / p.n_tiles) * CONV_BM;
      const int m = m0 + row;
      const bool mvalid = m < p.M_total;
      const __half* resp = has_res && mvalid ? p.res + static_cast<size_t>(m) * p.res_ld + p.res_coff + n0 : nullptr;
      uint32_t ra[16], rb[16];
      uint4 q[CONV_RES_DEPTH][2];
#pragma unroll
      for (int j = 0; j < CONV_RES_DEPTH; ++j) {       // residual of the first chunks: in flight while the MMAs still run
        q[j][0] = make_uint4(0, 0, 0, 0); q[j][1] = make_uint4(0, 0, 0, 0);
        if (resp && j < n_chunks) {
          q[j][0] = __ldg(reinterpret_cast<const uint4*>(resp + j * 16));
          q[j][1] = __ldg(reinterpret_cast<const uint4*>(resp + j * 16) + 1);
        }
      }
      mbar_wait(&acc_full[buf], (lt >> 1) & 1, 14);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.bn_tile);
      __syncwarp();
      tmem_ld_32x16(taddr, ra);
      for (int c0 = 0; c0 < n_chunks; c0 += CONV_RES_DEPTH) {
#pragma unroll
        for (int j = 0; j < CONV_RES_DEPTH; ++j) {
          const int c = c0 + j;
          if (c >= n_chunks) break;                    // warp-uniform
          // chunk c sits in ra (j even) / rb (j odd); start the TMEM load of chunk c+1 into the other set
          if (j & 1) tmem_ld_wait(rb); else tmem_ld_wait(ra);
          __syncwarp();
          if (c + 1 < n_chunks) {
            if (j & 1) tmem_ld_32x16(taddr + static_cast<uint32_t>((c + 1) * 16), ra);
            else tmem_ld_32x16(taddr + static_cast<uint32_t>((c + 1) * 16), rb);
          }
          if (mvalid) conv_finish_chunk(p, (j & 1) ? rb : ra, q[j], s_bias, m, n0 + c * 16, relu, has_res, out_f32);
          if (resp && c + CONV_RES_DEPTH < n_chunks) {  // refill this slot with the residual of chunk c + depth
            q[j][0] = __ldg(reinterpret_cast<const uint4*>(resp + (c + CONV_RES_DEPTH) * 16));
            q[j][1] = __ldg(reinterpret_cast<const uint4*>(resp + (c + CONV_RES_DEPTH) * 16) + 1);
          }
        }
      }
      __syncwarp();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  } else if (!p.tma_a) {
    // ---------------------------------------------------------------- A gather producers (4 warps)
    const int g = threadIdx.x - 192;                // 0..127
    const int chunk = g & 7, rbase = g >> 3;        // 8 lanes cover one 128-byte row
    const int HoWo = p.Ho * p.Wo;
    const bool use_l1 = p.flags & CF_GATHER_L1;
    if (p.pdl) pdl_wait();
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles) * CONV_BM;
      int base_off[8];
      short ih0[8], iw0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + rbase + 16 * i;
        if (m < p.M_total) {
          const int n = m / HoWo, rem = m - n * HoWo;
          const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
          ih0[i] = static_cast<short>(ho * p.stride - p.pad_h);
          iw0[i] = static_cast<short>(wo * p.stride - p.pad_w);
          base_off[i] = ((n * p.H + ih0[i]) * p.W + iw0[i]) * p.in_ld + p.in_coff;
        } else {
          ih0[i] = -30000; iw0[i] = -30000; base_off[i] = 0;      // always out of bounds -> zero fill
        }
      }
      for (int kb = 0; kb < p.nkb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1, 13);
        const int k = kb * 64 + chunk * 8;
        const int tap = k / p.cin, c = k - tap * p.cin;
        const int r = tap / p.kw, sx = tap - r * p.kw;
        const bool kvalid = k < p.k_real;
        const int tap_off = (r * p.W + sx) * p.in_ld + c;
        const uint32_t a_s = smem_u32(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ih = ih0[i] + r, iw = iw0[i] + sx;
          const bool ok = kvalid && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
          const __half* src = ok ? p.in + (base_off[i] + tap_off) : p.in;
          if (use_l1) cp_async_16_ca(a_s + sw128_offset(rbase + 16 * i, chunk), src, ok);
          else cp_async_16(a_s + sw128_offset(rbase + 16 * i, chunk), src, ok);
        }
        cp_async_mbar_arrive_noinc(&full[s]);         // counted arrival fires when this thread's copies have landed
      }
    }
    cp_async_wait_all();                              // do not exit with copies in flight
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace fire
