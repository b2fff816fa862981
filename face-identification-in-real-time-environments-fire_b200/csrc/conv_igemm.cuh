// conv_igemm.cuh - implicit-GEMM convolution on tcgen05 (K2 of DESIGN.md).
//
// One launch = one Conv(+folded BN)(+bias)(+residual)(+ReLU) layer of the FaceNet plan
// (reference: the onnxruntime session at facenet_gpu.py:127; graph in SURVEY App. A).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]      m = output pixel (b, ho, wo), n = output channel,
//                                          k = (tap r,s ; input channel c), tap-major
//   A is never materialised: each CTA builds its 128 x 64 K-block of A in shared memory
//     - TMA mode   (1x1, stride 1): the activation matrix [M, C] is a plain 2-D tensor -> one
//                   cp.async.bulk.tensor per K-block,
//     - gather mode (k x k / strided / padded): 128 producer threads issue 16-byte cp.async
//                   copies (zero-fill for padding / K tail) straight into the 128-byte-swizzled
//                   layout the UMMA descriptor expects,
//   W K-blocks always arrive by TMA.  One thread issues tcgen05.mma (M=128, N=bn_tile, K=16) into a
//   TMEM accumulator; four epilogue warps read it back (tcgen05.ld), add bias / residual, apply
//   ReLU, convert to fp16 (or keep fp32) and store at a channel offset of the destination
//   buffer, which is how concat costs nothing.
#pragma once

#include "fire_common.cuh"

namespace fire {

constexpr int CONV_BM = 128;
constexpr int CONV_THREADS = 192;          // warp0: TMA, warp1: MMA + TMEM owner, warps 2..5: gather + epilogue
constexpr int CONV_A_STAGE_BYTES = CONV_BM * 128;
constexpr int CONV_LAG = 2;                // cp.async groups in flight before a stage is signalled

constexpr int CF_RELU = 1, CF_RESIDUAL = 2, CF_OUT_F32 = 4;

struct ConvParams {
  const __half* in;  int in_ld, in_coff;
  void* out;         int out_ld, out_coff;
  const __half* res; int res_ld, res_coff;
  const float* bias;
  int H, W, Ho, Wo, kh, kw, stride, pad_h, pad_w;
  int cin, cout, k_real, nkb, flags, bn_tile, M_total, stages, tma_a, tmem_cols;
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__global__ void __launch_bounds__(CONV_THREADS)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.bn_tile * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.stages) * CONV_A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(p.stages) * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * CONV_BM;
  const int n0 = blockIdx.y * p.bn_tile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    if (p.tma_a) tma_prefetch_desc(&tmap_a);
  }
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t full_count = p.tma_a ? 1u : 1u + 128u;
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(b_stage_bytes) + (p.tma_a ? CONV_A_STAGE_BYTES : 0);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1, 11);
        mbar_arrive_expect_tx(&full[s], tx);
        tma_load_2d_hint(sB + static_cast<size_t>(s) * b_stage_bytes, &tmap_w, &full[s], kb * 64, n0, kEvictLast);
        if (p.tma_a) tma_load_2d(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES, &tmap_a, &full[s], kb * 64, m0);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(CONV_BM, p.bn_tile);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(&full[s], ph, 12);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + static_cast<size_t>(s) * CONV_A_STAGE_BYTES);
        const uint32_t b0 = smem_u32(sB + static_cast<size_t>(s) * b_stage_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int g = threadIdx.x - 64;                 // 0..127
    if (!p.tma_a) {
      // -------------------------------------------------------------- A gather producer
      const int chunk = g & 7, rbase = g >> 3;      // 8 lanes cover one 128-byte row
      int base_off[8];
      short ih0[8], iw0[8];
      const int HoWo = p.Ho * p.Wo;
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
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
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
          cp_async_16(a_s + sw128_offset(rbase + 16 * i, chunk), src, ok);
        }
        cp_async_commit();
        if (kb >= CONV_LAG) {
          cp_async_wait<CONV_LAG>();
          fence_proxy_async_smem();
          mbar_arrive(&full[(kb - CONV_LAG) % p.stages]);
        }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int kb = max(0, p.nkb - CONV_LAG); kb < p.nkb; ++kb) mbar_arrive(&full[kb % p.stages]);
    }
    // ---------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int m = m0 + row;
    const bool mvalid = m < p.M_total;
    mbar_wait(acc_full, 0, 14);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool relu = p.flags & CF_RELU, has_res = p.flags & CF_RESIDUAL, out_f32 = p.flags & CF_OUT_F32;
    for (int c0 = 0; c0 < p.bn_tile; c0 += 16) {
      uint32_t r[16];
      __syncwarp();                                   // tcgen05.ld is warp-collective: reconverge first
      tmem_ld_32x16(taddr + static_cast<uint32_t>(c0), r);
      tmem_ld_wait();
      if (!mvalid) continue;
      const int n = n0 + c0;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + __ldg(p.bias + n + j);
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + static_cast<size_t>(m) * p.res_ld + p.res_coff + n);
        uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1);
        const __half2* h0 = reinterpret_cast<const __half2*>(&q0);
        const __half2* h1 = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace fire
