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
// Flat mode (wide images, where R whole rows would fill little of the 128-row MMA): the OUTPUT buffer is allocated
// with row pitch Wbox, a tile is 128 consecutive positions of that pitched space (any start, mid-row included - the
// A descriptor just starts (f0 mod Wbox) rows into the patch) and the garbage columns land in the padding, which no
// consumer reads (their tensor maps end at Wo).
#pragma once

#include <type_traits>

#include "conv_igemm.cuh"

namespace fire {

constexpr int STRIP_MMA_WARPS = 2;                           // warps 1 and 11
constexpr int STRIP_THREADS = 32 * (3 + CONV_EPI_WARPS + STRIP_MMA_WARPS - 1);   // warp 0 producer, warps 1 / 11 MMA, warps 2-9 epilogue, warp 10 TMA stores
constexpr int STRIP_MAX_ACC = 8;                             // TMEM accumulators (tiles in flight between MMA and epilogue)

struct StripSmem {
  uint32_t a, b, bias, ones, zero, out, bars, total;
};
// The patch ring comes LAST and its stages are only as aligned as the operand swizzle needs (1024 B for 128-byte rows, 256 B for the
// 64- / 32-byte rows of the Cin = 32 / 16 layers: the swizzle is a function of ABSOLUTE shared-memory address bits, for the TMA's writes
// as for the UMMA's reads, so a stage may start at any multiple of the row size): Conv2d_2b gets six stages instead of four that way.
__host__ __device__ inline StripSmem strip_smem_layout(int stages, int a_stage_bytes, int nkb, int cout) {
  StripSmem L;
  uint32_t o = 0;
  L.b = o;    o += static_cast<uint32_t>(nkb) * cout * 128;
  L.bias = o; o += static_cast<uint32_t>(cout) * 16;
  L.ones = o; o += CONV_BM * 16;
  L.zero = o; o += static_cast<uint32_t>(cout > CONV_BM ? cout : CONV_BM) * 16;   // K-chunk 1 of the ones [128] and bias [cout] operands
  o = (o + 1023) & ~1023u;
  L.out = o;  o += 2u * CONV_BM * cout * 2;                  // [2 buffers][cout / box_cols boxes][128 rows][box_cols * 2 bytes]
  L.bars = o; o += 512;
  if ((a_stage_bytes & 1023) == 0) o = (o + 1023) & ~1023u;
  L.a = o;    o += static_cast<uint32_t>(stages) * a_stage_bytes;
  L.total = o;
  return L;
}

struct StripParams {
  const uint4* bias16;
  int cin, cout, kh, kw, pad_h, pad_w;
  int k16_steps;          // kh * kw * cin / 16
  int nkb;                // resident weight K-blocks (k_pad / 64)
  int Wbox, R, Hbox;      // patch: Wbox = Wo + kw - 1 pixels wide, R output rows per tile, Hbox = R + kh - 1
  int row_blocks;         // tiles per image: ceil(Ho / R), or ceil(Ho * Wbox / 128) in flat mode
  int total_tiles;        // images * row_blocks
  int flat;               // 1: a tile is 128 consecutive positions of the image's [Ho][Wbox] pitched position space (the output
                          //    buffer has row pitch Wbox, so the garbage columns land in its padding); 0: R whole rows
  int Ho;
  int a_stage_bytes, stages, tmem_cols, flags, pdl, box_cols;
  int n_mma_warps;        // MMA issuing warps (1 or 2): tile i of an interleaved group belongs to warp i % n_mma_warps.  One thread
                          // issues a tcgen05.mma every ~110 cycles whatever its size; issuers in different warps overlap
                          // (tools/umma_probe.cu part 6: N <= 64, 4 accumulators: 120 / 77 cycles per MMA with 1 / 2 threads).
                          // Measured: Conv2d_2a 100 -> 77 us, Conv2d_2b 112 -> 92 us with two; four warps were slower (118 / 121 us)
  int pair;               // 1: CTA pair (cluster of 2, tcgen05 cta_group::2): the pair takes the same position block of two
                          // consecutive images; ONE M = 256 tcgen05.mma of the leader covers both tiles, each CTA holds half of the
                          // weight rows.  A tcgen05.mma costs its issuing thread ~110 cycles whatever its size (probe part 7)
  int n_acc;              // TMEM accumulators (power of two, <= STRIP_MAX_ACC); the MMA warp interleaves n_acc / 2 tiles
  long long* trace;
  FastDiv d_rowblocks, d_wbox;
};

// kPair: CTA-pair instantiation (must be launched as clusters of 2: it contains cta_group::2 instructions, and a kernel that
// does cannot be launched without a cluster - "cluster misconfiguration")
template <bool kPair>
__global__ void __launch_bounds__(STRIP_THREADS, 1)
conv_strip_kernel_t(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_out, const StripParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const StripSmem L = strip_smem_layout(p.stages, p.a_stage_bytes, p.nkb, p.cout);
  uint8_t* sA = smem + L.a;
  uint8_t* sB = smem + L.b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + p.stages;
  uint64_t* acc_full = a_empty + p.stages;            // [STRIP_MAX_ACC]
  uint64_t* acc_empty = acc_full + STRIP_MAX_ACC;     // [STRIP_MAX_ACC]
  uint64_t* out_full = acc_empty + STRIP_MAX_ACC;     // [2] staging buffer written by the epilogue warps
  uint64_t* out_empty = out_full + 2;                 // [2] staging buffer read by the TMA store
  uint64_t* w_full = out_empty + 2;
  uint64_t* peer_full = w_full + 1;                   // [stages] pair mode, leader: the peer CTA's patch of this stage has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + p.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool pair = kPair;
  const uint32_t rank = pair ? cluster_ctarank() : 0u;              // 0 = leader (issues the MMAs of the pair)
  const int t_first = pair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int t_step = pair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // tile t = (image or image pair q, position block tin); this CTA's image:
  auto tile_image = [&](int q) { return pair ? 2 * q + static_cast<int>(rank) : q; };
  const int b_rows = pair ? p.cout / 2 : p.cout;                    // weight rows held by this CTA
  const int b_kb_bytes = b_rows * 128;
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
      // one arrival per epilogue WARP (256 per-thread arrivals on one barrier word serialise: ~250 cycles per barrier)
      for (int b = 0; b < STRIP_MAX_ACC; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], pair ? CONV_EPI_WARPS : CONV_EPI_WARPS / 2); }
      for (int s = 0; s < p.stages; ++s) mbar_init(&peer_full[s], 1);
      for (int b = 0; b < 2; ++b) { mbar_init(&out_full[b], CONV_EPI_WARPS / 2); mbar_init(&out_empty[b], 1); }
      mbar_init(w_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    if (!pair) tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  if (warp >= CONV_FIRST_EPI_WARP && warp < CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    const int t = threadIdx.x - CONV_FIRST_EPI_WARP * 32;                 // 0..255
    uint4* s_bias = reinterpret_cast<uint4*>(smem + L.bias);
    for (int i = t; i < b_rows; i += CONV_EPI_WARPS * 32) s_bias[i] = __ldg(p.bias16 + rank * b_rows + i);
    if (t < CONV_BM) reinterpret_cast<uint4*>(smem + L.ones)[t] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    if (t < (p.cout > CONV_BM ? p.cout : CONV_BM)) reinterpret_cast<uint4*>(smem + L.zero)[t] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (pair) {                                       // barriers of both CTAs are initialised before anything remote touches them
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
  if (p.pdl) pdl_launch_dependents();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: resident weights, then one patch per tile
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(p.nkb * b_kb_bytes));
        for (int kb = 0; kb < p.nkb; ++kb)
          tma_load_2d_hint(sB + static_cast<size_t>(kb) * b_kb_bytes, &tmap_w, w_full, kb * 64, static_cast<int>(rank) * b_rows, kEvictLast);
      }
      __syncwarp();
      if (p.pdl) pdl_wait();                                    // activations come from the previous layer
      const uint32_t box_bytes = static_cast<uint32_t>(in_row_bytes * p.Wbox * p.Hbox);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t_first; tile < p.total_tiles; tile += t_step) {
        const int q = fdiv(tile, p.d_rowblocks), tin = tile - q * p.row_blocks, n = tile_image(q);
        const int y0 = p.flat ? fdiv(tin * CONV_BM, p.d_wbox) : tin * p.R;      // first output row of the tile
        mbar_wait(&a_empty[s], ph ^ 1, 21);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[s], box_bytes);
          tma_load_4d(sA + static_cast<size_t>(s) * p.a_stage_bytes, &tmap_a, &a_full[s], 0, -p.pad_w, y0 - p.pad_h, n);
        }
        __syncwarp();
        if (lane == 0 && tile == t_first) CONV_TRACE(2);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || (warp > CONV_FIRST_EPI_WARP + CONV_EPI_WARPS && warp - (CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) < p.n_mma_warps)) {
    // ---------------------------------------------------------------- MMA issuers
    // `mw` (which tiles of a group this warp owns) is a COMPILE-TIME constant of each instantiation: predicates derived
    // from threadIdx are not provably warp-uniform, and ptxas then wraps every tcgen05.mma in its serialisation loop again.
    auto mma_role = [&](auto n_mw_c, auto mw_c) {
    constexpr int n_mw = decltype(n_mw_c)::value, mw = decltype(mw_c)::value;
    // The WHOLE warp runs the loop (so every address/descriptor stays in uniform registers) and one elected lane
    // issues: a divergent `if (lane == 0)` body makes ptxas wrap every tcgen05.mma in an ELECT + 7 x R2UR +
    // BRA.U.ANY serialisation loop, which costs ~100-200 cycles per MMA (tools/umma_probe.cu part 5/6).
    {
      const uint32_t idesc = umma_idesc_f16(pair ? 2 * CONV_BM : CONV_BM, p.cout);
      const uint32_t ones_addr = smem_u32(smem + L.ones), zero_addr = smem_u32(smem + L.zero), bias_addr = smem_u32(smem + L.bias);
      const uint64_t ones_desc = umma_desc_nosw(ones_addr, zero_addr - ones_addr, 128);
      const uint64_t bias_desc = umma_desc_nosw(bias_addr, zero_addr - bias_addr, 128);
      const uint32_t b_base = smem_u32(sB);
      const uint64_t a_desc_hi = umma_desc_swz(0u, static_cast<uint32_t>(in_row_bytes));     // everything but the address field
      const uint64_t b_desc_hi = umma_desc_sw128(0u);
      mbar_wait(w_full, 0, 22);
      tc_fence_after();
      // A small-N tcgen05.mma costs ~100 cycles when it depends on the previous one through the accumulator
      // (tools/umma_probe.cu part 5/6), so G = n_acc / 2 tiles are accumulated side by side: the K loop is the outer
      // loop and the G independent accumulators the inner one.
      const int G = p.n_acc >> 1;
      const int my_tiles = t_first < p.total_tiles ? (p.total_tiles - t_first + t_step - 1) / t_step : 0;
      int s = 0;
      uint32_t ph = 0;
      long long tm[3] = {0, 0, 0};
      const bool prof = p.trace != nullptr && (p.flags & CF_DBG_PHASES);
      for (int lt0 = 0; lt0 < my_tiles; lt0 += G) {
        const int cnt = min(G, my_tiles - lt0);
        long long c0 = prof ? clock64() : 0, c1;
        if (rank != 0) {
          // peer CTA of a pair: no MMAs here; tell the leader when this CTA's patches have landed
          int si = s;
          uint32_t phi = ph;
          for (int i = 0; i < cnt; ++i) {
            if (i % n_mw == mw) {
              mbar_wait(&a_full[si], phi, 24);
              if (lane == 0) mbar_arrive_remote_relaxed(&peer_full[si], 0u);
              __syncwarp();
            }
            if (++si == p.stages) { si = 0; phi ^= 1; }
          }
          for (int i = 0; i < cnt; ++i)
            if (++s == p.stages) { s = 0; ph ^= 1; }
          continue;
        }
        for (int i = mw; i < cnt; i += n_mw) {                  // this warp's tiles of the group: i = mw (mod n_mw)
          const int l = lt0 + i;
          mbar_wait(&acc_empty[l & (p.n_acc - 1)], (static_cast<uint32_t>(l / p.n_acc) & 1u) ^ 1u, 23);
        }
        tc_fence_after();
        if (prof) { c1 = clock64(); tm[0] += c1 - c0; c0 = c1; }
        if (elect_one()) {
          for (int i = mw; i < cnt; i += n_mw) {                // D = ones * bias^T
            const uint32_t dacc = tmem_base + static_cast<uint32_t>(((lt0 + i) & (p.n_acc - 1)) * p.cout);
            if (pair) umma_f16_2cta(dacc, ones_desc, bias_desc, idesc, 0u); else umma_f16(dacc, ones_desc, bias_desc, idesc, 0u);
          }
        }
        {
          int si = s;
          uint32_t phi = ph;
          for (int i = 0; i < cnt; ++i) {
            if (i % n_mw == mw) {
              mbar_wait(&a_full[si], phi, 24);
              if (pair) mbar_wait(&peer_full[si], phi, 28);
            }
            if (++si == p.stages) { si = 0; phi ^= 1; }
          }
        }
        tc_fence_after();
        if (prof) { c1 = clock64(); tm[1] += c1 - c0; c0 = c1; }
        if (lt0 == 0 && lane == 0 && mw == 0) CONV_TRACE(3);
        if (elect_one()) {
          const uint32_t a_ring = smem_u32(sA);
          uint32_t tile_base[STRIP_MAX_ACC / 2], tile_acc[STRIP_MAX_ACC / 2];   // per tile of the group: A start address, accumulator
          {
            int si = s;
#pragma unroll
            for (int i = 0; i < STRIP_MAX_ACC / 2; ++i) {
              const int tile_i = t_first + (lt0 + i) * t_step;
              const int tin = tile_i - fdiv(tile_i, p.d_rowblocks) * p.row_blocks;
              const int f0 = tin * CONV_BM;
              // flat mode: the tile starts (f0 mod Wbox) rows into its patch
              const uint32_t start = p.flat ? static_cast<uint32_t>((f0 - fdiv(f0, p.d_wbox) * p.Wbox) * in_row_bytes) : 0u;
              tile_base[i] = a_ring + static_cast<uint32_t>(si * p.a_stage_bytes) + start;
              tile_acc[i] = tmem_base + static_cast<uint32_t>(((lt0 + i) & (p.n_acc - 1)) * p.cout);
              if (++si == p.stages) si = 0;
            }
          }
          uint32_t row_off = 0, tap_off = 0, b_addr = b_base;     // byte offset of tap row r / of tap (r, sx) inside a patch
          int sx = 0, c0 = 0;
          for (int j = 0; j < p.k16_steps; ++j) {
            const uint64_t bdesc = b_desc_hi | static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4);
#pragma unroll
            for (int i = 0; i < STRIP_MAX_ACC / 2; ++i) {
              if (i < cnt && i % n_mw == mw) {
                const uint32_t a_addr = tile_base[i] + tap_off + static_cast<uint32_t>(c0 * 2);
                if (pair) umma_f16_2cta(tile_acc[i], a_desc_hi | static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4), bdesc, idesc, 1u);
                else umma_f16(tile_acc[i], a_desc_hi | static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4), bdesc, idesc, 1u);
              }
            }
            c0 += 16;
            b_addr += ((j & 3) == 3) ? static_cast<uint32_t>(b_kb_bytes - 96) : 32u;
            if (c0 == p.cin) {
              c0 = 0;
              tap_off += in_row_bytes;
              if (++sx == p.kw) { sx = 0; row_off += p.Wbox * in_row_bytes; tap_off = row_off; }
            }
          }
          int si = s;
          for (int i = 0; i < cnt; ++i) {
            if (i % n_mw == mw) {
              if (pair) { umma_commit_2cta(&a_empty[si]); umma_commit_2cta(&acc_full[(lt0 + i) & (p.n_acc - 1)]); }   // both CTAs
              else { umma_commit(&a_empty[si]); umma_commit(&acc_full[(lt0 + i) & (p.n_acc - 1)]); }
            }
            if (++si == p.stages) si = 0;
          }
        }
        __syncwarp();
        if (prof) { c1 = clock64(); tm[2] += c1 - c0; }
        for (int i = 0; i < cnt; ++i)
          if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (prof && lane == 0 && mw == 0) {
        long long* q = p.trace + 8 * 148 + blockIdx.x * 8;
        q[0] = tm[0]; q[1] = tm[1]; q[2] = tm[2]; q[7] = my_tiles;
      }
      if (lane == 0 && mw == 0) CONV_TRACE(4);
    }
    };
    using std::integral_constant;
    const int mwi = warp == 1 ? 0 : warp - (CONV_FIRST_EPI_WARP + CONV_EPI_WARPS);
    if (p.n_mma_warps == 1) mma_role(integral_constant<int, 1>{}, integral_constant<int, 0>{});
    else if (mwi == 0) mma_role(integral_constant<int, 2>{}, integral_constant<int, 0>{});
    else mma_role(integral_constant<int, 2>{}, integral_constant<int, 1>{});
  } else if (warp < CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue: two groups of 4 warps (one per TMEM lane quarter);
    // group g takes the tiles with lt & 1 == g and owns staging buffer g, so the fixed per-tile latencies (two barrier
    // waits, TMEM load, fences) of consecutive tiles overlap
    const int quarter = warp & 3, group = (warp - CONV_FIRST_EPI_WARP) >> 2;
    const bool relu = p.flags & CF_RELU;
    const int rowbytes = p.box_cols * 2, chunks_per_box = p.box_cols >> 4;
    const int box_bytes = CONV_BM * rowbytes;                   // one column box of the whole 128-row tile
    const int n_chunks = p.cout >> 4;
    const int m = quarter * 32 + lane;
    const uint32_t swz = p.box_cols == 64 ? (m & 7) : p.box_cols == 32 ? ((m >> 1) & 3) : ((m >> 2) & 1);
    const bool active = p.flat || quarter * 32 < p.R * p.Wbox;  // this quarter holds at least one real patch position
    const uint32_t stage_buf_bytes = static_cast<uint32_t>(CONV_BM * p.cout * 2);
    const uint32_t stage0 = smem_u32(smem + L.out);
    long long te[3] = {0, 0, 0};
    const bool prof = p.trace != nullptr && (p.flags & CF_DBG_PHASES) && threadIdx.x == CONV_FIRST_EPI_WARP * 32;     // group 0's tiles only
    int lt = 0;
    for (int tile = t_first; tile < p.total_tiles; tile += t_step, ++lt) {
      if ((lt & 1) != group) continue;
      const int buf = lt & (p.n_acc - 1), ob = lt & 1;
      const uint32_t stage = stage0 + static_cast<uint32_t>(ob) * stage_buf_bytes;
      long long c0 = prof ? clock64() : 0, c1;
      mbar_wait(&out_empty[ob], ((lt >> 1) & 1) ^ 1, 26);       // the store of two tiles ago has left this staging buffer
      if (prof) { c1 = clock64(); te[0] += c1 - c0; c0 = c1; }
      mbar_wait(&acc_full[buf], static_cast<uint32_t>(lt / p.n_acc) & 1u, 25);
      tc_fence_after();
      if (prof) { c1 = clock64(); te[1] += c1 - c0; c0 = c1; }
      if (lt == 0 && threadIdx.x == CONV_FIRST_EPI_WARP * 32) CONV_TRACE(5);
      if (active) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.cout);
        const uint32_t my_row = stage + static_cast<uint32_t>(m * rowbytes);
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t ra[16];
          __syncwarp();
          tmem_ld_32x16(taddr + static_cast<uint32_t>(c * 16), ra);
          tmem_ld_wait(ra);
          const int box = c / chunks_per_box, cb = c - box * chunks_per_box;
          conv_stage_chunk(ra, relu, my_row + static_cast<uint32_t>(box * box_bytes), static_cast<uint32_t>(2 * cb), swz);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();                                 // staging writes -> visible to the TMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&out_full[ob]);
        if (rank != 0) mbar_arrive_remote_relaxed(&acc_empty[buf], 0u); else mbar_arrive(&acc_empty[buf]);     // the leader's MMA warps own the accumulators
      }
      if (prof) { c1 = clock64(); te[2] += c1 - c0; }
    }
    if (prof) {
      long long* q = p.trace + 8 * 148 + blockIdx.x * 8;
      q[3] = te[0]; q[4] = te[1]; q[5] = te[2];
    }
    if (threadIdx.x == CONV_FIRST_EPI_WARP * 32) CONV_TRACE(6);
  } else if (warp == CONV_FIRST_EPI_WARP + CONV_EPI_WARPS) {
    // ---------------------------------------------------------------- TMA store warp: one 4-D store per column box of a tile
    const int rowbytes = p.box_cols * 2, box_bytes = CONV_BM * rowbytes, n_boxes = p.cout / p.box_cols;
    const uint32_t stage_buf_bytes = static_cast<uint32_t>(CONV_BM * p.cout * 2);
    const uint32_t stage0 = smem_u32(smem + L.out);
    if (p.pdl) pdl_wait();                                      // output writes must not overtake readers of the previous layers
    int lt = 0;
    for (int tile = t_first; tile < p.total_tiles; tile += t_step, ++lt) {
      const int ob = lt & 1;
      const int q = fdiv(tile, p.d_rowblocks), tin = tile - q * p.row_blocks, n = tile_image(q);
      mbar_wait(&out_full[ob], (lt >> 1) & 1, 27);
      if (elect_one()) {
        if (!(p.flags & CF_DBG_NOSTORE)) {
          for (int b = 0; b < n_boxes; ++b) {
            const uint32_t src = stage0 + static_cast<uint32_t>(ob) * stage_buf_bytes + static_cast<uint32_t>(b * box_bytes);
            if (p.flat) tma_store_3d(&tmap_out, src, b * p.box_cols, tin * CONV_BM, n);      // {channel, position, image}
            else tma_store_4d(&tmap_out, src, b * p.box_cols, 0, tin * p.R, n);               // {channel, x, y, image}
          }
        }
        bulk_commit_group();
        bulk_wait_read_all();                                   // staging buffer read: hand it back (keeping a store in flight
        mbar_arrive(&out_empty[ob]);                            // measured slower: it queues behind the patch loads)
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_all();                           // stores complete before the CTA exits
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();                     // the peer's accumulators and barriers are still in use until both are done
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc2_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols)); else tmem_dealloc_rt(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  if (threadIdx.x == 0) CONV_TRACE(7);
}

}  // namespace fire
