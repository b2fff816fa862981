// conv_strip.cuh - stride-1 k x k convolution WITHOUT im2col: halo patch + row-shifted UMMA descriptors.
//
// Same layer contract as conv_igemm.cuh (Conv + folded BN + bias + ReLU, NHWC fp16 in/out, channel-offset
// stores), used for the stride-1 layers whose weights fit in shared memory (Conv2d_2a/2b, the 3x3 convs of
// Block35; reference graph: facenet_gpu.py:127).
//
// A tile is R consecutive output rows of one image.  ONE 4-D TMA load brings the input patch
// [R + kh - 1 rows][Wbox = Wo + kw - 1 pixels][Cin] into shared memory (out-of-image coordinates are zero-filled by
// the TMA: that is the 'same' padding), as Hbox * Wbox consecutive swizzled rows of Cin * 2 bytes.  GEMM row m is
// patch position (m / Wbox, m % Wbox), and the A operand of tap (r, s) is simply the SAME patch with the descriptor
// start address advanced by (r * Wbox + s) rows - tcgen05.mma applies the swizzle to absolute shared-memory address
// bits, so any row shift is legal (measured on the B200: tools/umma_probe.cu part 2).  Every input byte crosses
// L2 -> SM once per tile instead of kh * kw times, there are no gather threads and no per-K-block handshakes:
// per tile the producer issues one TMA, the MMA thread issues kh*kw*Cin/16 tcgen05.mma, the epilogue converts.
// The Wbox - Wo garbage columns of each row (and rows past Ho) are computed but never stored: the 4-D TMA store
// clips them against the output tensor's extents.  Weights stay resident for the whole launch.
#pragma once

#include "conv_igemm.cuh"

namespace fire {

constexpr int STRIP_THREADS = 32 * (2 + CONV_EPI_WARPS);     // warp 0 producer, warp 1 MMA, warps 2-9 epilogue

struct StripSmem {
  uint32_t a, b, bias, ones, zero, out, bars, total;
};
__host__ __device__ inline StripSmem strip_smem_layout(int stages, int a_stage_bytes, int nkb, int cout) {
  StripSmem L;
  uint32_t o = 0;
  L.a = o;    o += static_cast<uint32_t>(stages) * a_stage_bytes;
  L.b = o;    o += static_cast<uint32_t>(nkb) * cout * 128;
  L.bias = o; o += static_cast<uint32_t>(cout) * 16;
  L.ones = o; o += CONV_BM * 16;
  L.zero = o; o += 256 * 16;
  o = (o + 1023) & ~1023u;
  L.out = o;  o += 2u * CONV_BM * cout * 2;                  // [2 buffers][cout / box_cols boxes][128 rows][box_cols * 2 bytes]
  L.bars = o; o += 256;
  L.total = o;
  return L;
}

struct StripParams {
  const uint4* bias16;
  int cin, cout, kh, kw, pad_h, pad_w;
  int k16_steps;          // kh * kw * cin / 16
  int nkb;                // resident weight K-blocks (k_pad / 64)
  int Wbox, R, Hbox;      // patch: Wbox = Wo + kw - 1 pixels wide, R output rows per tile, Hbox = R + kh - 1
  int row_blocks;         // ceil(Ho / R)
  int total_tiles;        // images * row_blocks
  int a_stage_bytes, stages, tmem_cols, flags, pdl, box_cols;
  long long* trace;
  FastDiv d_rowblocks;
};

__global__ void __launch_bounds__(STRIP_THREADS, 1)
conv_strip_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_out, const StripParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const StripSmem L = strip_smem_layout(p.stages, p.a_stage_bytes, p.nkb, p.cout);
  uint8_t* sA = smem + L.a;
  uint8_t* sB = smem + L.b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + p.stages;
  uint64_t* acc_full = a_empty + p.stages;    // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_kb_bytes = p.cout * 128;
  const int in_row_bytes = p.cin * 2;
  if (threadIdx.x == 0) CONV_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], CONV_EPI_WARPS * 32); }
      mbar_init(w_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  if (warp >= CONV_FIRST_EPI_WARP) {
    const int t = threadIdx.x - CONV_FIRST_EPI_WARP * 32;                 // 0..255
    uint4* s_bias = reinterpret_cast<uint4*>(smem + L.bias);
    for (int i = t; i < p.cout; i += CONV_EPI_WARPS * 32) s_bias[i] = __ldg(p.bias16 + i);
    if (t < CONV_BM) reinterpret_cast<uint4*>(smem + L.ones)[t] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(smem + L.zero)[t] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CONV_TRACE(1);
  if (p.pdl) pdl_launch_dependents();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: resident weights, then one patch per tile
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(p.nkb * b_kb_bytes));
        for (int kb = 0; kb < p.nkb; ++kb)
          tma_load_2d_hint(sB + static_cast<size_t>(kb) * b_kb_bytes, &tmap_w, w_full, kb * 64, 0, kEvictLast);
      }
      __syncwarp();
      if (p.pdl) pdl_wait();                                    // activations come from the previous layer
      const uint32_t box_bytes = static_cast<uint32_t>(in_row_bytes * p.Wbox * p.Hbox);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = fdiv(tile, p.d_rowblocks), y0 = (tile - n * p.row_blocks) * p.R;
        mbar_wait(&a_empty[s], ph ^ 1, 21);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[s], box_bytes);
          tma_load_4d(sA + static_cast<size_t>(s) * p.a_stage_bytes, &tmap_a, &a_full[s], 0, -p.pad_w, y0 - p.pad_h, n);
        }
        __syncwarp();
        if (lane == 0 && tile == static_cast<int>(blockIdx.x)) CONV_TRACE(2);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // The WHOLE warp runs the loop (so every address/descriptor stays in uniform registers) and one elected lane
    // issues: a divergent `if (lane == 0)` body makes ptxas wrap every tcgen05.mma in an ELECT + 7 x R2UR +
    // BRA.U.ANY serialisation loop, which costs ~100-200 cycles per MMA (tools/umma_probe.cu part 5/6).
    {
      const uint32_t idesc = umma_idesc_f16(CONV_BM, p.cout);
      const uint32_t ones_addr = smem_u32(smem + L.ones), zero_addr = smem_u32(smem + L.zero), bias_addr = smem_u32(smem + L.bias);
      const uint64_t ones_desc = umma_desc_nosw(ones_addr, zero_addr - ones_addr, 128);
      const uint64_t bias_desc = umma_desc_nosw(bias_addr, zero_addr - bias_addr, 128);
      const uint32_t b_base = smem_u32(sB);
      const uint64_t a_desc_hi = umma_desc_swz(0u, static_cast<uint32_t>(in_row_bytes));     // everything but the address field
      const uint64_t b_desc_hi = umma_desc_sw128(0u);
      mbar_wait(w_full, 0, 22);
      tc_fence_after();
      int lt = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1, 23);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * p.cout);
        if (elect_one()) umma_f16(d, ones_desc, bias_desc, idesc, 0u);           // D = ones * bias^T
        mbar_wait(&a_full[s], ph, 24);
        tc_fence_after();
        if (lt == 0 && lane == 0) CONV_TRACE(3);
        const uint32_t a_base = smem_u32(sA + static_cast<size_t>(s) * p.a_stage_bytes);
        if (elect_one()) {
          uint32_t a_row = a_base, a_tap = a_base, b_addr = b_base;           // start of tap row r / of tap (r, sx)
          int sx = 0, c0 = 0;
          for (int j = 0; j < p.k16_steps; ++j) {
            umma_f16(d, a_desc_hi | static_cast<uint64_t>(((a_tap + c0 * 2) & 0x3FFFF) >> 4),
                     b_desc_hi | static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4), idesc, 1u);
            c0 += 16;
            b_addr += ((j & 3) == 3) ? static_cast<uint32_t>(b_kb_bytes - 96) : 32u;
            if (c0 == p.cin) {
              c0 = 0;
              a_tap += in_row_bytes;
              if (++sx == p.kw) { sx = 0; a_row += p.Wbox * in_row_bytes; a_tap = a_row; }
            }
          }
          umma_commit(&a_empty[s]);
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (lane == 0) CONV_TRACE(4);
    }
  } else {
    // ---------------------------------------------------------------- epilogue (8 warps, 2 per TMEM lane quarter)
    const int quarter = warp & 3, half = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const bool relu = p.flags & CF_RELU;
    const bool leader = threadIdx.x == CONV_FIRST_EPI_WARP * 32;
    const int rowbytes = p.box_cols * 2, chunks_per_box = p.box_cols >> 4;
    const int box_bytes = CONV_BM * rowbytes;                   // one column box of the whole 128-row tile
    const int n_boxes = p.cout / p.box_cols, n_chunks = p.cout >> 4;
    const int m = quarter * 32 + lane;
    const uint32_t swz = p.box_cols == 64 ? (m & 7) : p.box_cols == 32 ? ((m >> 1) & 3) : ((m >> 2) & 1);
    const bool active = quarter * 32 < p.R * p.Wbox;            // this quarter holds at least one real patch position
    const uint32_t stage_buf_bytes = static_cast<uint32_t>(CONV_BM * p.cout * 2);
    const uint32_t stage0 = smem_u32(smem + L.out);
    uint32_t sbuf = 0;
    if (p.pdl && leader) pdl_wait();
    long long acc_t[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool prof = leader && p.trace != nullptr;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      long long c0 = prof ? clock64() : 0, c1;
      const int n = fdiv(tile, p.d_rowblocks), y0 = (tile - n * p.row_blocks) * p.R;
      const uint32_t stage = stage0 + sbuf * stage_buf_bytes;
      sbuf ^= 1;
      if (leader) bulk_wait_read_1();                           // the store issued two tiles ago has left this staging buffer
      if (prof) { c1 = clock64(); acc_t[0] += c1 - c0; c0 = c1; }
      named_bar_sync(5, CONV_EPI_WARPS * 32);
      if (prof) { c1 = clock64(); acc_t[1] += c1 - c0; c0 = c1; }
      mbar_wait(&acc_full[buf], (lt >> 1) & 1, 25);
      tc_fence_after();
      if (prof) { c1 = clock64(); acc_t[2] += c1 - c0; c0 = c1; }
      if (lt == 0 && leader) CONV_TRACE(5);
      if (active) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.cout);
        const uint32_t my_row = stage + static_cast<uint32_t>(m * rowbytes);
        for (int c = half; c < n_chunks; c += 2) {
          uint32_t ra[16];
          __syncwarp();
          tmem_ld_32x16(taddr + static_cast<uint32_t>(c * 16), ra);
          tmem_ld_wait(ra);
          const int box = c / chunks_per_box, cb = c - box * chunks_per_box;
          conv_stage_chunk(ra, relu, my_row + static_cast<uint32_t>(box * box_bytes), static_cast<uint32_t>(2 * cb), swz);
        }
      }
      __syncwarp();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
      if (prof) { c1 = clock64(); acc_t[3] += c1 - c0; c0 = c1; }
      fence_proxy_async_smem();
      if (prof) { c1 = clock64(); acc_t[4] += c1 - c0; c0 = c1; }
      named_bar_sync(5, CONV_EPI_WARPS * 32);
      if (prof) { c1 = clock64(); acc_t[5] += c1 - c0; c0 = c1; }
      if (leader && !(p.flags & CF_DBG_NOSTORE)) {
        for (int b = 0; b < n_boxes; ++b)
          tma_store_4d(&tmap_out, stage + static_cast<uint32_t>(b * box_bytes), b * p.box_cols, 0, y0, n);
        bulk_commit_group();
      }
      if (prof) { c1 = clock64(); acc_t[6] += c1 - c0; }
    }
    if (leader) { bulk_wait_all(); CONV_TRACE(6); }
    if (prof) {
      for (int i = 0; i < 7; ++i) p.trace[8 * 256 * 0 + 8 * 148 + blockIdx.x * 8 + i] = acc_t[i];
      p.trace[8 * 148 + blockIdx.x * 8 + 7] = lt;
    }
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
