// knn.cu - exact cosine top-k over the enrolled gallery (K3 of DESIGN.md).
//
// Replaces hnswlib's Index.add_items / knn_query as used by modules/hnsw_manager.py:127,137,147,237
// with a brute-force search whose ids equal hnswlib BFIndex's:
//
//   knn_add_kernel      rows -> x * 1/(||x||+1e-30)  -> fp32 master copy + fp16 operand copy
//   knn_prep_kernel     same normalisation for the query batch
//   knn_scan_kernel     tcgen05 GEMM  S = Q16 . G16^T  (fp16 operands, fp32 accumulate in TMEM);
//                       the epilogue warps read S out of TMEM and keep a per-query sorted list of the
//                       KP best (score, row) in REGISTERS - the score matrix is never written anywhere.
//   knn_rerank_kernel   merges the per-split lists, recomputes the KP survivors exactly in fp32 from
//                       the master copy, orders by (distance asc, id asc) and PROVES the fp16 filter
//                       could not have dropped a true top-k row (|fp16 score - exact| <= eps); queries
//                       for which that proof fails are queued for
//   knn_exact_*         a plain fp32 scan of the whole shard (rare; correctness backstop).
//   knn_merge_kernel    multi-GPU: merge G per-shard top-k lists after the all-gather.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <cstdlib>
#include <new>

#include "fire_common.cuh"
#include "fire_internal.h"

namespace fire {

constexpr int KNN_BM = 128;                  // queries per CTA  (UMMA M)
constexpr int KNN_BN = 256;                  // gallery rows per accumulator tile (UMMA N)
constexpr int KNN_ISSUERS = 3;               // max TMA issuing threads: one thread sustains only ~1 box per 600-800 cycles (tools/umma_probe.cu)
constexpr int KNN_THREADS = 192 + 32 * (KNN_ISSUERS - 1);   // warp0 TMA, warp1 MMA, warps 2..5 epilogue, warps 6.. extra TMA issuers
constexpr int KNN_A_KB_BYTES = KNN_BM * 128; // one 64-wide K block of the query tile
constexpr int KNN_B_STAGE_BYTES = KNN_BN * 128;
constexpr int KNN_MAX_LISTS_PER_LANE = 10;   // merge routines handle up to 320 sorted lists
constexpr float KNN_DEFAULT_EPS = 3e-5f;     // slack on top of the measured fp16 rounding-error bound (fp32 accumulation order)
constexpr size_t KNN_SMEM_BUDGET = 232448 - 2048;

struct KnnScanParams {
  int nkb;              // D / 64
  int n_rows;           // rows in this shard
  int S;                // gallery splits
  int rows_per_split;   // multiple of KNN_BN
  int QB;               // query blocks of 128
  int stages;
  int n_issuers;        // divides `stages`, so a ring slot always belongs to the same issuing thread (no parity aliasing)
  float* cand_score;    // [QB*128][S][KP]
  uint32_t* cand_idx;
};

// ------------------------------------------------------------------------------------------------
// normalise rows like hnswlib's cosine space: x * (1 / (sqrt(sum x^2) + 1e-30))
// one warp per row; writes the fp32 normalised row and its fp16 rounding.
// Also measures ||x^ - fp16(x^)||_2 of every row: per query into row_err[], max over gallery rows into *max_err
// (positive floats order like their bit patterns, so atomicMax on the bits is a float max).
// `prenormalized`: the rows are stored gallery rows (already x^), used as queries as they are (fire_knn_search_rows).
__global__ void knn_normalize_kernel(const float* __restrict__ in, size_t n, int D, float* __restrict__ out32,
                                     __half* __restrict__ out16, size_t n_pad16, float* __restrict__ row_err,
                                     int* __restrict__ max_err, int prenormalized) {
  size_t row = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= n_pad16) return;
  if (row >= n) {   // zero padding rows of the fp16 operand (query tile padding)
    for (int c = lane * 2; c < D; c += 64) *reinterpret_cast<uint32_t*>(out16 + row * D + c) = 0u;
    return;
  }
  const float* x = in + row * D;
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x + c);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float inv = prenormalized ? 1.0f : 1.0f / (sqrtf(s) + 1e-30f);
  float e2 = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x + c);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    *reinterpret_cast<float4*>(out32 + row * D + c) = v;
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    *reinterpret_cast<uint2*>(out16 + row * D + c) = pk;
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    const float d0 = v.x - fa.x, d1 = v.y - fa.y, d2 = v.z - fb.x, d3 = v.w - fb.y;
    e2 = fmaf(d0, d0, e2); e2 = fmaf(d1, d1, e2); e2 = fmaf(d2, d2, e2); e2 = fmaf(d3, d3, e2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  if (lane == 0) {
    const float e = sqrtf(e2) * 1.0001f;              // round the norm itself upwards
    if (row_err) row_err[row] = e;
    if (max_err) atomicMax(max_err, __float_as_int(e));
  }
}

// ------------------------------------------------------------------------------------------------
// Where results go.  Either two arrays (dist f32 [Q][k], ids i64 [Q][k]) or ONE packed array of 12-byte records
// {f32 distance, i64 id} stored as three 32-bit words - the message of the multi-GPU exchange, so that a single
// all-gather carries both.  id = row * id_stride + id_offset maps a shard row to its global label
// (contiguous shards: stride 1, offset = first row; interleaved shards: stride = world, offset = rank).
struct KnnOut {
  float* dist;
  long long* ids;
  int32_t* packed;
  long long id_offset, id_stride;
  __device__ __forceinline__ void put(size_t slot, float d, uint32_t row) const {
    const long long id = row == 0xFFFFFFFFu ? -1ll : static_cast<long long>(row) * id_stride + id_offset;
    if (packed) {
      packed[slot * 3] = __float_as_int(d);
      packed[slot * 3 + 1] = static_cast<int32_t>(static_cast<unsigned long long>(id) & 0xFFFFFFFFull);
      packed[slot * 3 + 2] = static_cast<int32_t>(static_cast<unsigned long long>(id) >> 32);
    } else {
      dist[slot] = d;
      ids[slot] = id;
    }
  }
};

// Padding fill: every slot = (FLT_MAX, -1).  Used when a shard holds no rows at all.
__global__ void knn_fill_padding_kernel(KnnOut out, size_t n_slots) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_slots; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out.put(i, FLT_MAX, 0xFFFFFFFFu);
}

// ------------------------------------------------------------------------------------------------
// Sorted (descending score) candidate list held in registers.
template <int KP>
struct RegList {
  float s[KP];
  uint32_t ix[KP];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KP; ++j) { s[j] = -FLT_MAX; ix[j] = 0xFFFFFFFFu; }
  }
  __device__ __forceinline__ float worst() const { return s[KP - 1]; }
  // precondition: v > s[KP-1].  Bubble the new entry up from the bottom.
  __device__ __forceinline__ void insert(float v, uint32_t id) {
    s[KP - 1] = v; ix[KP - 1] = id;
#pragma unroll
    for (int j = KP - 1; j > 0; --j) {
      bool sw = s[j] > s[j - 1];
      float ts = sw ? s[j - 1] : s[j];   uint32_t ti = sw ? ix[j - 1] : ix[j];
      s[j - 1] = sw ? s[j] : s[j - 1];   ix[j - 1] = sw ? ix[j] : ix[j - 1];
      s[j] = ts; ix[j] = ti;
    }
  }
};

// r[j] with a run-time j, as a 5-level select tree (keeps r[] in registers).
__device__ __forceinline__ float pick32(const uint32_t (&r)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return __uint_as_float((j & 16) ? d[1] : d[0]);
}

// ------------------------------------------------------------------------------------------------
// One CTA = one (query block, gallery split).  The 128 x D query tile stays resident in shared
// memory; gallery K-blocks stream through a TMA ring; two 256-column accumulators alternate in TMEM
// so the top-k epilogue of tile t overlaps the MMAs of tile t+1.
// kPair (clusters of two, QB even, tcgen05 cta_group::2): the two CTAs of a cluster hold consecutive query blocks and scan the
// SAME gallery split as ONE M = 256 MMA per K step.  Each CTA keeps its own 128 queries (A) and loads only HALF of every
// gallery tile's rows (its half of the B operand): every SM ingests half of the gallery bytes.  Only the leader (cluster rank
// 0) issues MMAs; completions are multicast to both CTAs' barriers; the peer's otherwise idle MMA warp relays "my operands
// have landed" to the leader (same protocol as conv_igemm_kernel_t<true>).  Bit-identical results (tests/test_gpu_knn.py), but
// NO gain measured on the B200 in any regime - like sharing the stages of two independent CTAs by multicast TMA, tried just
// before (round 2): the scan is not bound by gallery bytes per SM.  Kept behind FIRE_B200_KNN_PAIR=1 for that record only.
template <int KP, bool kPair>
__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                const KnnScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int b_stage_bytes = kPair ? KNN_B_STAGE_BYTES / 2 : KNN_B_STAGE_BYTES;
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.nkb) * KNN_A_KB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(p.stages) * b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* peer_a_full = tempty + 2;             // pair mode, leader: the peer's query tile / gallery half of stage s has landed
  uint64_t* peer_full = peer_a_full + 1;          // [stages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + p.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x % p.QB, split = blockIdx.x / p.QB;
  const int row0 = split * p.rows_per_split;
  const int row1 = min(row0 + p.rows_per_split, p.n_rows);
  const int n_tiles = row1 > row0 ? (row1 - row0 + KNN_BN - 1) / KNN_BN : 0;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(a_full, 1);
      mbar_init(peer_a_full, 1);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&peer_full[s], 1); }
      // one arrival per epilogue WARP (per-thread arrivals on one barrier word serialise); pair: both CTAs' warps release the leader's
      for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kPair ? 8 : 4); }
      fence_barrier_init();
    }
    __syncwarp();
    if (!kPair) tmem_alloc<512>(tmem_slot);
  }
  if (kPair) {                                      // barriers of both CTAs exist before anything remote touches them
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_alloc2_rt(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int issuer = warp == 0 ? 0 : (warp >= 6 ? warp - 5 : -1);
  if (issuer >= 0) {
    // ------------------------------------------------------------------ TMA producers (stage `it` belongs to issuer it % n_issuers)
    if (lane == 0 && issuer < p.n_issuers) {
      if (issuer == 0) {
        mbar_arrive_expect_tx(a_full, static_cast<uint32_t>(p.nkb) * KNN_A_KB_BYTES);
        for (int kb = 0; kb < p.nkb; ++kb)
          tma_load_2d_hint(sA + static_cast<size_t>(kb) * KNN_A_KB_BYTES, &tmap_q, a_full, kb * 64, qb * KNN_BM, kEvictLast);
      }
      int it = 0, s = 0, turn = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          if (turn == issuer) {
            mbar_wait(&empty[s], ph ^ 1, 1);
            mbar_arrive_expect_tx(&full[s], b_stage_bytes);
            // pair: this CTA's half of the tile's rows (the tensor map's box is 128 rows then)
            tma_load_2d(sB + static_cast<size_t>(s) * b_stage_bytes, &tmap_g, &full[s], kb * 64,
                        row0 + t * KNN_BN + (kPair ? static_cast<int>(crank) * (KNN_BN / 2) : 0));
          }
          if (++turn == p.n_issuers) turn = 0;
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread); the peer's thread is the relay
    if (lane == 0 && crank != 0) {
      mbar_wait(a_full, 0, 6);
      mbar_arrive_remote_relaxed(peer_a_full, 0u);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t)
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full[s], ph, 7);
          mbar_arrive_remote_relaxed(&peer_full[s], 0u);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kPair ? 2 * KNN_BM : KNN_BM, KNN_BN);
      mbar_wait(a_full, 0, 2);
      if (kPair) mbar_wait(peer_a_full, 0, 8);
      tc_fence_after();
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1, 3);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(buf * KNN_BN);
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full[s], ph, 4);
          if (kPair) mbar_wait(&peer_full[s], ph, 9);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + static_cast<size_t>(kb) * KNN_A_KB_BYTES);
          const uint32_t b0 = smem_u32(sB + static_cast<size_t>(s) * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kPair) umma_f16_2cta(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (kPair) umma_commit_2cta(&empty[s]); else umma_commit(&empty[s]);      // frees the stage (in both CTAs) once these MMAs have read it
        }
        if (kPair) umma_commit_2cta(&tfull[buf]); else umma_commit(&tfull[buf]);    // accumulator tile complete (each CTA reads its own 128 lanes)
      }
    }
  } else {
    // ------------------------------------------------------------------ top-k epilogue (4 warps)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int qrow = quarter * 32 + lane;         // query row inside the block
    RegList<KP> list;
    list.init();
    float thr = -FLT_MAX;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(&tfull[buf], (t >> 1) & 1, 5);
      tc_fence_after();
      const int col0 = row0 + t * KNN_BN;
      const int nvalid = min(KNN_BN, row1 - col0);
      // One 32-column chunk of this lane's query row: block maximum against the list's threshold (almost always the end of
      // it once the list has warmed up), else the few candidates above the threshold are inserted.
      auto consume = [&](uint32_t (&r)[32], int c) {
        const int lim = nvalid - c * 32;          // >= 32 for full chunks
        if (lim < 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j >= lim) r[j] = __float_as_uint(-FLT_MAX);
        }
        float m = __uint_as_float(r[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
        if (m > thr) {                             // rare once the list has warmed up
          uint32_t mask = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(r[j]) > thr) ? (1u << j) : 0u;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float v = pick32(r, j);
            if (v > thr) {
              list.insert(v, static_cast<uint32_t>(col0 + c * 32 + j));
              thr = list.worst();
            }
          }
        }
      };
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * KNN_BN);
      const int n_chunks = (nvalid + 31) >> 5;
      if (KP <= 16) {
        // tcgen05.ld has ~500 cycles of latency and the block maximum of 32 values costs ~40 instructions: walking the eight
        // chunks one after the other left the accumulator busy for ~4 us per tile - longer than the tile's MMAs (2.1 us), so
        // the EPILOGUE set the pace of the whole scan (round-2 ncu: tensor pipe 65 % at Q = 4096, 48 % at Q = 256).  The
        // load of chunk c+1 is now in flight while chunk c is reduced (two register buffers).
        uint32_t ra[32], rb[32];
        __syncwarp();
        tmem_ld_32x32(tbase, ra);
#pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          tmem_ld_wait(ra);
          __syncwarp();
          if (c + 1 < n_chunks) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 1) * 32), rb);
          consume(ra, c);
          if (c + 1 < n_chunks) {
            tmem_ld_wait(rb);
            __syncwarp();
            if (c + 2 < n_chunks) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 2) * 32), ra);
            consume(rb, c + 1);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          __syncwarp();                                // tcgen05.ld is warp-collective: reconverge first
          tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), r);
          tmem_ld_wait(r);
          consume(r, c);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (crank != 0) mbar_arrive_remote_relaxed(&tempty[buf], 0u); else mbar_arrive(&tempty[buf]); }
    }
    const size_t q = static_cast<size_t>(qb) * KNN_BM + qrow;
    float* os = p.cand_score + (q * p.S + split) * KP;
    uint32_t* oi = p.cand_idx + (q * p.S + split) * KP;
#pragma unroll
    for (int j = 0; j < KP; j += 4) {
      *reinterpret_cast<float4*>(os + j) = make_float4(list.s[j], list.s[j + 1], list.s[j + 2], list.s[j + 3]);
      *reinterpret_cast<uint4*>(oi + j) = make_uint4(list.ix[j], list.ix[j + 1], list.ix[j + 2], list.ix[j + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();                  // no CTA leaves while its peer can still signal its barriers or the pair's MMAs run
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2_rt(tmem_base, 512u); else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative k-way merge of G lists, each sorted ascending by (key, id).  Lane l owns lists
// l, l+32, ...  `get(list, pos, &key, &id)` reads one entry.  Emits the n_out smallest entries in
// order through `emit(rank, key, id)` (called by all lanes with identical arguments).
struct KeyId {
  float key;
  unsigned long long id;
};
__device__ __forceinline__ bool keyid_less(float ka, unsigned long long ia, float kb, unsigned long long ib) {
  return ka < kb || (ka == kb && ia < ib);
}

template <class Get, class Emit>
__device__ __forceinline__ void warp_merge_lists(int G, int len, int n_out, Get get, Emit emit) {
  const int lane = threadIdx.x & 31;
  int head[KNN_MAX_LISTS_PER_LANE];
  float hk[KNN_MAX_LISTS_PER_LANE];
  unsigned long long hid[KNN_MAX_LISTS_PER_LANE];
#pragma unroll
  for (int i = 0; i < KNN_MAX_LISTS_PER_LANE; ++i) {
    head[i] = 0;
    const int l = lane + 32 * i;
    hk[i] = FLT_MAX; hid[i] = ~0ull;
    if (l < G && len > 0) get(l, 0, hk[i], hid[i]);
  }
  for (int r = 0; r < n_out; ++r) {
    float bk = FLT_MAX; unsigned long long bid = ~0ull; int bi = -1;
#pragma unroll
    for (int i = 0; i < KNN_MAX_LISTS_PER_LANE; ++i) {
      if (lane + 32 * i < G && head[i] < len && (bi < 0 || keyid_less(hk[i], hid[i], bk, bid))) {
        bk = hk[i]; bid = hid[i]; bi = i;
      }
    }
    // warp arg-min over (key, id); lanes with bi < 0 carry (FLT_MAX, ~0) and lose every comparison
    float wk = bk; unsigned long long wid = bid; int wl = bi >= 0 ? lane : 64;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ok = __shfl_xor_sync(0xffffffffu, wk, o);
      unsigned long long oid = __shfl_xor_sync(0xffffffffu, wid, o);
      int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      bool take = (ol < 64) && (wl >= 64 || keyid_less(ok, oid, wk, wid) || (ok == wk && oid == wid && ol < wl));
      if (take) { wk = ok; wid = oid; wl = ol; }
    }
    emit(r, wk, wid, wl < 64);
    if (wl == lane && bi >= 0) {
#pragma unroll
      for (int i = 0; i < KNN_MAX_LISTS_PER_LANE; ++i) {
        if (i == bi) {
          head[i]++;
          hk[i] = FLT_MAX; hid[i] = ~0ull;
          if (head[i] < len) get(lane + 32 * i, head[i], hk[i], hid[i]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One warp per query: merge split lists by fp16 score, exact fp32 re-rank, order, prove or flag.
template <int KP>
__global__ void knn_rerank_kernel(const float* __restrict__ qn, const float* __restrict__ g32,
                                  const float* __restrict__ cand_score, const uint32_t* __restrict__ cand_idx, int S,
                                  int D, int Q, int k, float slack, const float* __restrict__ q_err,
                                  const int* __restrict__ g_max_err, KnnOut out, uint32_t* __restrict__ flagged,
                                  uint32_t* __restrict__ flagged_count, unsigned long long* __restrict__ stats) {
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] += static_cast<unsigned long long>(Q);      // queries answered by this handle
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  constexpr int EPL = (KP + 31) / 32;   // entries per lane
  float sel_score[EPL];
  uint32_t sel_idx[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) { sel_score[e] = -FLT_MAX; sel_idx[e] = 0xFFFFFFFFu; }

  const float* cs = cand_score + static_cast<size_t>(q) * S * KP;
  const uint32_t* ci = cand_idx + static_cast<size_t>(q) * S * KP;
  auto get = [&](int l, int pos, float& key, unsigned long long& id) {
    key = -cs[l * KP + pos];                       // descending score == ascending -score
    id = ci[l * KP + pos];
  };
  auto emit = [&](int r, float key, unsigned long long id, bool valid) {
    if ((r & 31) == lane) {
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        if (e == (r >> 5)) { sel_score[e] = valid ? -key : -FLT_MAX; sel_idx[e] = valid ? static_cast<uint32_t>(id) : 0xFFFFFFFFu; }
    }
  };
  warp_merge_lists(S, KP, KP, get, emit);

  // exact fp32 distances of the survivors: one candidate per round, the whole warp on one row (512 contiguous bytes per load).
  // (Round 2 tried four candidates per round with eight lanes each, and merging the lists from shared memory: the single-query
  // latency did not move - it is instruction latency of one warp, ~30 us - and the 4096-query batch got slower, 48 -> 66 us.)
  const float* qv = qn + static_cast<size_t>(q) * D;
  float dist[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) dist[e] = FLT_MAX;
  for (int r = 0; r < KP; ++r) {
    uint32_t idx = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e)
      if (e == (r >> 5)) idx = __shfl_sync(0xffffffffu, sel_idx[e], r & 31);
    if (idx == 0xFFFFFFFFu) continue;             // warp-uniform
    const float* gv = g32 + static_cast<size_t>(idx) * D;
    float acc = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
      float4 a = *reinterpret_cast<const float4*>(qv + c);
      float4 b = __ldg(reinterpret_cast<const float4*>(gv + c));
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((r & 31) == lane) {
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        if (e == (r >> 5)) { const float dd = 1.0f - acc; dist[e] = dd == dd ? dd : FLT_MAX; }    // a NaN (NaN/Inf in the query) sorts last instead of poisoning the ranking
    }
  }

  // rank by (distance asc, id asc); invalid entries (FLT_MAX, ~0) sort last and, being equal to one another, are told apart by
  // their list position - every entry gets a distinct rank, so all k output slots are always written (padding = (FLT_MAX, -1))
  int rank[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) rank[e] = 0;
  for (int r = 0; r < KP; ++r) {
    float od = 0.f; uint32_t oi = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e)
      if (e == (r >> 5)) { od = __shfl_sync(0xffffffffu, dist[e], r & 31); oi = __shfl_sync(0xffffffffu, sel_idx[e], r & 31); }
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int me = e * 32 + lane;
      if (me < KP && me != r && (keyid_less(od, oi, dist[e], sel_idx[e]) || (od == dist[e] && oi == sel_idx[e] && r < me))) rank[e]++;
    }
  }
  float kth_dist = -FLT_MAX;   // distance of rank k-1
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int me = e * 32 + lane;
    if (me < KP && rank[e] < k) {
      out.put(static_cast<size_t>(q) * k + rank[e], dist[e], sel_idx[e]);
      if (rank[e] == k - 1) kth_dist = dist[e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kth_dist = fmaxf(kth_dist, __shfl_xor_sync(0xffffffffu, kth_dist, o));
  // worst surviving fp16 score: entry KP-1 (valid only if the merged list is full)
  float cmin = __shfl_sync(0xffffffffu, sel_score[EPL - 1], (KP - 1) & 31);
  uint32_t last_idx = __shfl_sync(0xffffffffu, sel_idx[EPL - 1], (KP - 1) & 31);
  const bool list_full = last_idx != 0xFFFFFFFFu;
  // |<q,g> - <q16,g16>| <= ||dq||*||g16|| + ||q16||*||dg|| + ||dq||*||dg||  with the MEASURED rounding-error norms
  // (dq = q^ - fp16(q^), dg likewise, max over the shard) plus slack for the fp32 accumulation order.
  // Rows outside the list have fp16 score <= cmin, hence exact cosine <= cmin + eps: they cannot enter the
  // top-k iff cmin + eps < (1 - kth_dist).
  const float eq = q_err[q], eg = __int_as_float(*g_max_err);
  const float eps = (eq + eg) * 1.0005f + eq * eg + slack;
  if (lane == 0 && list_full && (cmin + eps >= 1.0f - kth_dist)) {
    uint32_t slot = atomicAdd(flagged_count, 1u);
    flagged[slot] = static_cast<uint32_t>(q);
  }
}

// ------------------------------------------------------------------------------------------------
// Flagged queries (the merged-list proof failed): exact fp32 work, bounded by what the scan already knows.
//
//   knn_refine_kernel   one block per flagged query: the exact distances of ALL S x KP candidates the scan kept (every
//                       split's own list, not only the merged KP best) give a new exact top-k.  A row that is in NO list
//                       was dropped inside its own split s, so its fp16 score is <= that split's KP-th score cmin_s: split s
//                       can still hide a true top-k row only if its list is full and cmin_s + eps >= 1 - kth.  Splits
//                       that pass need no further work (the usual case: one split's KP-th score is far below the global
//                       one).  Exactly one unsafe split -> work item (query, split): only that split's rows are scanned.
//                       Two or more -> work item (query, all rows).
//   knn_exact_scan_kernel   rows of every work item's range split over the blocks -> per-block partial top-k
//   knn_exact_merge_kernel  per work item: partial lists (+ the candidate list, duplicates dropped) -> output row
//   knn_exact_overflow_kernel  flagged queries beyond EXACT_CAP (pathological galleries): one block scans the whole
//                       shard for one query.
// All kernels read the counts on the device and return at once when there is nothing to do, so they are enqueued
// unconditionally (no host round trip).  counters: [0] flagged queries, [1] work items.
constexpr int EXACT_WARPS = 8;
constexpr int EXACT_CAP = 256;
constexpr int EXACT_MAX_BLOCKS = 296;        // partial lists per work item (+ 1 candidate list) <= 32 * KNN_MAX_LISTS_PER_LANE

struct ExactSmem {
  float q[512];
  float d[EXACT_WARPS][64];
  uint32_t i[EXACT_WARPS][64];
};
struct KnnWorkItem {
  uint32_t q;       // query row
  uint32_t f;       // index into flagged[] (= slot of the candidate list in ref_dist / ref_idx)
  int split;        // gallery split to scan, -1 = every row of the shard
  int pad;
};

__device__ __forceinline__ void exact_lists_reset(ExactSmem& sm, const float* __restrict__ qn, uint32_t q, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) sm.q[c] = qn[static_cast<size_t>(q) * D + c];
  for (int j = lane; j < 64; j += 32) { sm.d[warp][j] = FLT_MAX; sm.i[warp][j] = 0xFFFFFFFFu; }
  __syncthreads();
}
// this warp: exact distance of gallery row r, inserted into the warp's sorted list of the k best (distance asc, id asc)
__device__ __forceinline__ void exact_consume_row(ExactSmem& sm, const float* __restrict__ g32, int D, int k, uint32_t r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* gv = g32 + static_cast<size_t>(r) * D;
  float acc = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float4 b = __ldg(reinterpret_cast<const float4*>(gv + c));
    acc = fmaf(sm.q[c], b.x, acc); acc = fmaf(sm.q[c + 1], b.y, acc);
    acc = fmaf(sm.q[c + 2], b.z, acc); acc = fmaf(sm.q[c + 3], b.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  float d = 1.0f - acc;
  if (!(d == d)) d = FLT_MAX;
  if (lane == 0 && keyid_less(d, static_cast<unsigned long long>(r), sm.d[warp][k - 1], sm.i[warp][k - 1])) {
    int j = k - 1;
    while (j > 0 && keyid_less(d, static_cast<unsigned long long>(r), sm.d[warp][j - 1], sm.i[warp][j - 1])) {
      sm.d[warp][j] = sm.d[warp][j - 1]; sm.i[warp][j] = sm.i[warp][j - 1]; --j;
    }
    sm.d[warp][j] = d; sm.i[warp][j] = r;
  }
  __syncwarp();
}
// every warp scans rows r0+warp, r0+warp+8, ... < r1 and keeps its k best in sm.d/sm.i[warp]
__device__ __forceinline__ void exact_block_scan(ExactSmem& sm, const float* __restrict__ qn,
                                                 const float* __restrict__ g32, uint32_t q, int D, int k, int r0,
                                                 int r1) {
  const int warp = threadIdx.x >> 5;
  exact_lists_reset(sm, qn, q, D);
  for (int r = r0 + warp; r < r1; r += EXACT_WARPS) exact_consume_row(sm, g32, D, k, static_cast<uint32_t>(r));
  __syncthreads();
}

__global__ void __launch_bounds__(EXACT_WARPS * 32)
knn_refine_kernel(const float* __restrict__ qn, const float* __restrict__ g32, const float* __restrict__ cand_score,
                  const uint32_t* __restrict__ cand_idx, int S, int KP, int D, int k, float slack,
                  const float* __restrict__ q_err, const int* __restrict__ g_max_err, const uint32_t* __restrict__ flagged,
                  uint32_t* __restrict__ counters, float* __restrict__ ref_dist, uint32_t* __restrict__ ref_idx,
                  KnnWorkItem* __restrict__ work, KnnOut out, unsigned long long* __restrict__ stats, int Q) {
  __shared__ ExactSmem sm;
  __shared__ float s_kth;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[1] += static_cast<unsigned long long>(counters[0]);     // queries whose proof failed
  (void)Q;
  const uint32_t nf = min(counters[0], static_cast<uint32_t>(EXACT_CAP));
  for (uint32_t f = blockIdx.x; f < nf; f += gridDim.x) {
    const uint32_t q = flagged[f];
    const float* cs = cand_score + static_cast<size_t>(q) * S * KP;
    const uint32_t* ci = cand_idx + static_cast<size_t>(q) * S * KP;
    exact_lists_reset(sm, qn, q, D);
    const int total = S * KP;
    for (int c = warp; c < total; c += EXACT_WARPS) {
      const uint32_t idx = ci[c];
      if (idx != 0xFFFFFFFFu) exact_consume_row(sm, g32, D, k, idx);     // warp-uniform
    }
    __syncthreads();
    if (warp == 0) {
      auto get = [&](int l, int pos, float& key, unsigned long long& id) { key = sm.d[l][pos]; id = sm.i[l][pos]; };
      auto emit = [&](int r, float key, unsigned long long id, bool valid) {
        if (lane == 0) {
          const uint32_t row = valid ? static_cast<uint32_t>(id) : 0xFFFFFFFFu;
          const float d = valid && row != 0xFFFFFFFFu ? key : FLT_MAX;
          out.put(static_cast<size_t>(q) * k + r, d, row);
          ref_dist[static_cast<size_t>(f) * 64 + r] = d;
          ref_idx[static_cast<size_t>(f) * 64 + r] = row;
          if (r == k - 1) s_kth = d;
        }
      };
      warp_merge_lists(EXACT_WARPS, k, k, get, emit);
    }
    __syncthreads();
    if (warp == 0) {
      const float kth = s_kth;
      const float eq = q_err[q], eg = __int_as_float(*g_max_err);
      const float eps = (eq + eg) * 1.0005f + eq * eg + slack;
      int unsafe = 0, first = -1;
      for (int s0 = 0; s0 < S; s0 += 32) {
        const int sidx = s0 + lane;
        bool bad = false;
        if (sidx < S) {
          const bool full = ci[static_cast<size_t>(sidx) * KP + KP - 1] != 0xFFFFFFFFu;
          bad = full && (cs[static_cast<size_t>(sidx) * KP + KP - 1] + eps >= 1.0f - kth);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, bad);
        if (m && first < 0) first = s0 + __ffs(m) - 1;
        unsafe += __popc(m);
      }
      if (lane == 0 && unsafe > 0) {
        const uint32_t slot = atomicAdd(&counters[1], 1u);
        KnnWorkItem w;
        w.q = q; w.f = f; w.split = unsafe == 1 ? first : -1; w.pad = 0;
        work[slot] = w;
        atomicAdd(&stats[unsafe == 1 ? 2 : 3], 1ull);
      }
    }
  }
}

__global__ void __launch_bounds__(EXACT_WARPS * 32)
knn_exact_scan_kernel(const float* __restrict__ qn, const float* __restrict__ g32, int n_rows, int rows_per_split, int D, int k,
                      const KnnWorkItem* __restrict__ work, const uint32_t* __restrict__ counters,
                      float* __restrict__ part_dist, uint32_t* __restrict__ part_idx) {
  __shared__ ExactSmem sm;
  const int lane = threadIdx.x & 31;
  const uint32_t nw = min(counters[1], static_cast<uint32_t>(EXACT_CAP));
  for (uint32_t w = 0; w < nw; ++w) {
    const KnnWorkItem it = work[w];
    const int R0 = it.split < 0 ? 0 : it.split * rows_per_split;
    const int R1 = it.split < 0 ? n_rows : min(n_rows, R0 + rows_per_split);
    const int rows_per_block = (R1 - R0 + gridDim.x - 1) / gridDim.x;
    const int r0 = min(R1, R0 + static_cast<int>(blockIdx.x) * rows_per_block), r1 = min(R1, r0 + rows_per_block);
    exact_block_scan(sm, qn, g32, it.q, D, k, r0, r1);
    if (threadIdx.x < 32) {
      float* od = part_dist + (static_cast<size_t>(w) * gridDim.x + blockIdx.x) * k;
      uint32_t* oi = part_idx + (static_cast<size_t>(w) * gridDim.x + blockIdx.x) * k;
      auto get = [&](int l, int pos, float& key, unsigned long long& id) { key = sm.d[l][pos]; id = sm.i[l][pos]; };
      auto emit = [&](int r, float key, unsigned long long id, bool valid) {
        if (lane == 0) { od[r] = valid ? key : FLT_MAX; oi[r] = valid ? static_cast<uint32_t>(id) : 0xFFFFFFFFu; }
      };
      warp_merge_lists(EXACT_WARPS, k, k, get, emit);
    }
  }
}

// one warp per work item: merge the per-block partial lists (and, for a single-split item, the candidate list of the other
// splits - rows of the scanned split appear in both, the second copy is dropped) and overwrite the query's output row
__global__ void knn_exact_merge_kernel(const float* __restrict__ part_dist, const uint32_t* __restrict__ part_idx, int nblk, int k,
                                       const KnnWorkItem* __restrict__ work, const uint32_t* __restrict__ counters,
                                       const float* __restrict__ ref_dist, const uint32_t* __restrict__ ref_idx, KnnOut out) {
  const uint32_t nw = min(counters[1], static_cast<uint32_t>(EXACT_CAP));
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (uint32_t w = blockIdx.x * wpb + (threadIdx.x >> 5); w < nw; w += gridDim.x * wpb) {
    const KnnWorkItem it = work[w];
    const float* pd = part_dist + static_cast<size_t>(w) * nblk * k;
    const uint32_t* pi = part_idx + static_cast<size_t>(w) * nblk * k;
    const float* rd = ref_dist + static_cast<size_t>(it.f) * 64;
    const uint32_t* ri = ref_idx + static_cast<size_t>(it.f) * 64;
    const int lists = it.split < 0 ? nblk : nblk + 1;
    auto get = [&](int l, int pos, float& key, unsigned long long& id) {
      if (l < nblk) { key = pd[l * k + pos]; id = pi[l * k + pos]; }
      else { key = rd[pos]; id = ri[pos]; }
    };
    int cnt = 0;
    unsigned long long prev = ~0ull;
    auto emit = [&](int r, float key, unsigned long long id, bool valid) {
      (void)r;
      const bool real = valid && id != 0xFFFFFFFFull;
      if (cnt < k && (!real || id != prev)) {             // all lanes see the same stream: cnt / prev stay warp-uniform
        if (lane == 0) out.put(static_cast<size_t>(it.q) * k + cnt, real ? key : FLT_MAX, real ? static_cast<uint32_t>(id) : 0xFFFFFFFFu);
        ++cnt;
      }
      if (real) prev = id;
    };
    warp_merge_lists(lists, k, 2 * k, get, emit);
  }
}

__global__ void __launch_bounds__(EXACT_WARPS * 32)
knn_exact_overflow_kernel(const float* __restrict__ qn, const float* __restrict__ g32, int n_rows, int D, int k,
                          const uint32_t* __restrict__ flagged, const uint32_t* __restrict__ counters, KnnOut out) {
  __shared__ ExactSmem sm;
  const int lane = threadIdx.x & 31;
  const uint32_t total = counters[0];
  for (uint32_t f = EXACT_CAP + blockIdx.x; f < total; f += gridDim.x) {
    const uint32_t q = flagged[f];
    exact_block_scan(sm, qn, g32, q, D, k, 0, n_rows);
    if (threadIdx.x < 32) {
      auto get = [&](int l, int pos, float& key, unsigned long long& id) { key = sm.d[l][pos]; id = sm.i[l][pos]; };
      auto emit = [&](int r, float key, unsigned long long id, bool valid) {
        if (lane == 0) {
          const bool real = valid && id != 0xFFFFFFFFull;
          out.put(static_cast<size_t>(q) * k + r, real ? key : FLT_MAX, real ? static_cast<uint32_t>(id) : 0xFFFFFFFFu);
        }
      };
      warp_merge_lists(EXACT_WARPS, k, k, get, emit);
    }
  }
}

// multi-GPU merge: G per-shard lists per query (each ascending) -> global top-k per query.  The lists come either as two
// arrays dists/ids [G][Q][k] or as ONE array of packed 12-byte records [G][Q][k][3] (what a single all-gather delivers).
__global__ void knn_merge_kernel(const float* __restrict__ dists, const long long* __restrict__ ids, const int32_t* __restrict__ packed,
                                 int Q, int k, int G, float* __restrict__ out_dist, long long* __restrict__ out_ids) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  auto get = [&](int l, int pos, float& key, unsigned long long& id) {
    const size_t o = (static_cast<size_t>(l) * Q + q) * k + pos;
    if (packed) {
      key = __int_as_float(packed[o * 3]);
      id = static_cast<unsigned long long>(static_cast<uint32_t>(packed[o * 3 + 1])) |
           (static_cast<unsigned long long>(static_cast<uint32_t>(packed[o * 3 + 2])) << 32);
    } else {
      key = dists[o];
      id = static_cast<unsigned long long>(ids[o]);
    }
  };
  auto emit = [&](int r, float key, unsigned long long id, bool valid) {
    if (lane == 0) {
      out_dist[static_cast<size_t>(q) * k + r] = valid ? key : FLT_MAX;
      out_ids[static_cast<size_t>(q) * k + r] = valid ? static_cast<long long>(id) : -1ll;
    }
  };
  warp_merge_lists(G, k, k, get, emit);
}

}  // namespace fire

// ================================================================================================
// host side
// ================================================================================================
using namespace fire;

struct fire_knn {
  int D = 0;
  int device = 0;
  size_t capacity = 0, count = 0;
  float* g32 = nullptr;       // [capacity][D] normalised fp32 master
  __half* g16 = nullptr;      // [capacity][D] fp16 operand copy
  // search scratch (grow-only)
  int q_cap = 0;              // queries the scratch can hold (multiple of 128)
  int s_cap = 0, kp_cap = 0;
  float* qn32 = nullptr;
  float* q_err = nullptr;       // [q_cap] ||q^ - fp16(q^)||
  int* g_max_err = nullptr;     // [1] max over gallery rows of ||g^ - fp16(g^)|| (float bits)
  __half* q16 = nullptr;
  float* cand_score = nullptr;
  uint32_t* cand_idx = nullptr;
  uint32_t* flagged = nullptr;       // [q_cap]
  uint32_t* counters = nullptr;      // [4]: flagged queries, work items
  unsigned long long* stats = nullptr;   // device: {queries_total, queries_flagged, single-split scans, whole-shard scans}
  float* part_dist = nullptr;
  uint32_t* part_idx = nullptr;
  float* ref_dist = nullptr;         // [EXACT_CAP][64] exact top-k over the candidates of every flagged query
  uint32_t* ref_idx = nullptr;
  KnnWorkItem* work = nullptr;       // [EXACT_CAP]
  int exact_blocks = 0;
  float eps = KNN_DEFAULT_EPS;
  // the last search's decomposition, for a deferred fallback (host path: launched only when a query was flagged)
  int last_S = 0, last_KP = 0, last_k = 0, last_Q = 0, last_rows_per_split = 0, last_n_rows = 0;
  uint32_t* host_flagged = nullptr;  // pinned [1]
  // staging for *_host calls
  float* stage_q = nullptr; float* stage_d = nullptr; long long* stage_i = nullptr;
  size_t stage_q_cap = 0, stage_o_cap = 0;
  float* stage_rows = nullptr; size_t stage_rows_cap = 0;
  float* dev_q = nullptr; float* dev_d = nullptr; long long* dev_i = nullptr;
  size_t dev_q_cap = 0, dev_o_cap = 0;
};

static int knn_ensure_scratch(fire_knn* h, int Q, int S, int KP) {
  const int q_pad = (Q + KNN_BM - 1) / KNN_BM * KNN_BM;
  if (q_pad > h->q_cap || S > h->s_cap || KP > h->kp_cap) {
    const int nq = std::max(q_pad, h->q_cap), ns = std::max(S, h->s_cap), nkp = std::max(KP, h->kp_cap);
    FIRE_CUDA(cudaDeviceSynchronize());
    cudaFree(h->qn32); cudaFree(h->q16); cudaFree(h->cand_score); cudaFree(h->cand_idx); cudaFree(h->flagged); cudaFree(h->q_err);
    h->q_err = nullptr; h->qn32 = nullptr; h->q16 = nullptr; h->cand_score = nullptr; h->cand_idx = nullptr; h->flagged = nullptr;
    h->q_cap = 0; h->s_cap = 0; h->kp_cap = 0;
    FIRE_CUDA(cudaMalloc(&h->qn32, sizeof(float) * nq * h->D));
    FIRE_CUDA(cudaMalloc(&h->q_err, sizeof(float) * nq));
    FIRE_CUDA(cudaMalloc(&h->q16, sizeof(__half) * nq * h->D));
    FIRE_CUDA(cudaMalloc(&h->cand_score, sizeof(float) * static_cast<size_t>(nq) * ns * nkp));
    FIRE_CUDA(cudaMalloc(&h->cand_idx, sizeof(uint32_t) * static_cast<size_t>(nq) * ns * nkp));
    FIRE_CUDA(cudaMalloc(&h->flagged, sizeof(uint32_t) * nq));
    h->q_cap = nq; h->s_cap = ns; h->kp_cap = nkp;
  }
  if (!h->counters) {
    FIRE_CUDA(cudaMalloc(&h->counters, sizeof(uint32_t) * 4));
    FIRE_CUDA(cudaMalloc(&h->stats, sizeof(unsigned long long) * 4));
    FIRE_CUDA(cudaMemset(h->stats, 0, sizeof(unsigned long long) * 4));
    h->exact_blocks = std::min(EXACT_MAX_BLOCKS, 2 * device_sm_count());
  }
  if (!h->part_dist) {
    FIRE_CUDA(cudaMalloc(&h->part_dist, sizeof(float) * EXACT_CAP * h->exact_blocks * 64));
    FIRE_CUDA(cudaMalloc(&h->part_idx, sizeof(uint32_t) * EXACT_CAP * h->exact_blocks * 64));
    FIRE_CUDA(cudaMalloc(&h->ref_dist, sizeof(float) * EXACT_CAP * 64));
    FIRE_CUDA(cudaMalloc(&h->ref_idx, sizeof(uint32_t) * EXACT_CAP * 64));
    FIRE_CUDA(cudaMalloc(&h->work, sizeof(KnnWorkItem) * EXACT_CAP));
  }
  return FIRE_OK;
}

template <int KP, bool kPair>
static int knn_launch_scan(fire_knn* h, const CUtensorMap& tq, const CUtensorMap& tg, const KnnScanParams& p,
                           size_t smem_bytes, cudaStream_t st) {
  static bool attr_done[FIRE_MAX_DEVICES] = {};        // the opt-in is per device (context), not per process
  if (!attr_done[h->device]) {
    FIRE_CUDA(cudaFuncSetAttribute(knn_scan_kernel<KP, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(KNN_SMEM_BUDGET + 1024)));
    attr_done[h->device] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.QB * p.S));
  cfg.blockDim = dim3(KNN_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kPair ? 1 : 0;
  FIRE_CUDA(cudaLaunchKernelEx(&cfg, knn_scan_kernel<KP, kPair>, tq, tg, p));
  count_launch();
  return FIRE_OK;
}

// The whole search: normalise the queries, tensor-core scan, exact re-rank with proof, bounded exact fallback.
// `stored_first` >= 0: the queries are the stored (already normalised) rows [stored_first, stored_first + Q).
// `allow_short`: k may exceed the number of stored rows; missing entries come back as (FLT_MAX, -1) (shards of a sharded gallery).
// refine -> exact scan -> merge -> overflow for the queries the rerank flagged (all four return at once when there are none)
static int knn_launch_fallback(fire_knn* h, const KnnOut& out, cudaStream_t st) {
  const int sms = device_sm_count();
  knn_refine_kernel<<<sms, EXACT_WARPS * 32, 0, st>>>(h->qn32, h->g32, h->cand_score, h->cand_idx, h->last_S, h->last_KP, h->D, h->last_k, h->eps, h->q_err,
                                                      h->g_max_err, h->flagged, h->counters, h->ref_dist, h->ref_idx, h->work, out, h->stats, h->last_Q);
  knn_exact_scan_kernel<<<h->exact_blocks, EXACT_WARPS * 32, 0, st>>>(h->qn32, h->g32, h->last_n_rows, h->last_rows_per_split, h->D, h->last_k, h->work,
                                                                      h->counters, h->part_dist, h->part_idx);
  knn_exact_merge_kernel<<<32, 256, 0, st>>>(h->part_dist, h->part_idx, h->exact_blocks, h->last_k, h->work, h->counters, h->ref_dist, h->ref_idx, out);
  knn_exact_overflow_kernel<<<h->exact_blocks, EXACT_WARPS * 32, 0, st>>>(h->qn32, h->g32, h->last_n_rows, h->D, h->last_k, h->flagged, h->counters, out);
  FIRE_LAUNCH_CHECK("knn exact fallback");
  count_launch(4);
  return FIRE_OK;
}

// `defer_fallback`: the caller will look at the flagged count on the host (it synchronises anyway) and launch the fallback
// only when needed - four launches less on the latency path of a single query.
static int knn_search_impl(fire_knn* h, const float* queries, long long stored_first, int Q, int k, const KnnOut& out, bool allow_short,
                           cudaStream_t st, bool defer_fallback = false) {
  if (Q <= 0) return fail(FIRE_ERR_ARG, "fire_knn_search: Q=%d", Q);
  if (k < 1 || k > 64) return fail(FIRE_ERR_UNSUPPORTED, "fire_knn_search: k=%d outside [1,64]", k);
  if (static_cast<size_t>(k) > h->count && !allow_short)
    return fail(FIRE_ERR_STATE, "fire_knn_search: k=%d exceeds the %zu stored rows", k, h->count);
  if (h->count == 0) {                              // empty shard: nothing but padding
    const size_t n = static_cast<size_t>(Q) * k;
    knn_fill_padding_kernel<<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 1024)), 256, 0, st>>>(out, n);
    FIRE_LAUNCH_CHECK("knn_fill_padding_kernel");
    count_launch();
    h->last_Q = 0;
    return FIRE_OK;
  }
  const int D = h->D, nkb = D / 64;
  const int KP = k <= 10 ? 16 : 64;
  const int n_rows = static_cast<int>(h->count);
  const int QB = (Q + KNN_BM - 1) / KNN_BM;
  const int tiles_total = (n_rows + KNN_BN - 1) / KNN_BN;
  const int sms = device_sm_count();
  // work decomposition: QB x S CTAs, about W waves over the SMs, >= ~200 tiles per CTA when possible
  long long w = (static_cast<long long>(tiles_total) * QB) / (static_cast<long long>(sms) * 200);
  const int W = static_cast<int>(std::max<long long>(1, std::min<long long>(8, w)));
  int S = std::max(1, (W * sms) / QB);
  S = std::min(S, std::min(tiles_total, 32 * KNN_MAX_LISTS_PER_LANE));
  const int tiles_per_split = (tiles_total + S - 1) / S;
  S = (tiles_total + tiles_per_split - 1) / tiles_per_split;

  int rc = knn_ensure_scratch(h, Q, S, KP);
  if (rc != FIRE_OK) return rc;

  // normalised queries (fp32 for the exact re-rank, fp16 tile-padded for the tensor cores)
  {
    const size_t q_pad = static_cast<size_t>(QB) * KNN_BM;
    const size_t blocks = (q_pad + 7) / 8;
    const float* src = stored_first >= 0 ? h->g32 + static_cast<size_t>(stored_first) * D : queries;
    knn_normalize_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, static_cast<size_t>(Q), D, h->qn32, h->q16, q_pad, h->q_err,
                                                                         nullptr, stored_first >= 0 ? 1 : 0);
    FIRE_LAUNCH_CHECK("knn_normalize_kernel(query)");
    count_launch();
  }
  CUtensorMap tq, tg;
  rc = make_tmap_f16_2d(&tq, h->q16, static_cast<uint64_t>(QB) * KNN_BM, D, static_cast<uint64_t>(D) * 2, KNN_BM);
  if (rc != FIRE_OK) return rc;
  // CTA pairs (see knn_scan_kernel) are OFF by default: measured on the B200 (profiles/r02_knn_experiments.txt) they change
  // neither the few-query-blocks regime (Q = 256: 0.325 vs 0.327 ms) nor the large batches (Q = 4096: 3.40 vs 3.53 ms)
  bool pair = false;
  if (const char* e = getenv("FIRE_B200_KNN_PAIR")) pair = QB % 2 == 0 && e[0] == '1';        // A/B experiments, parity test
  rc = make_tmap_f16_2d(&tg, h->g16, static_cast<uint64_t>(n_rows), D, static_cast<uint64_t>(D) * 2, pair ? KNN_BN / 2 : KNN_BN);
  if (rc != FIRE_OK) return rc;

  KnnScanParams p;
  p.nkb = nkb; p.n_rows = n_rows; p.S = S; p.rows_per_split = tiles_per_split * KNN_BN; p.QB = QB;
  const size_t a_bytes = static_cast<size_t>(nkb) * KNN_A_KB_BYTES;
  const size_t bar_bytes = 512;
  const size_t b_stage = pair ? KNN_B_STAGE_BYTES / 2 : KNN_B_STAGE_BYTES;
  int stages = static_cast<int>((KNN_SMEM_BUDGET - a_bytes - bar_bytes) / b_stage);
  stages = std::max(2, std::min(stages, pair ? 12 : 6));
  p.stages = stages;
  p.n_issuers = stages % 3 == 0 ? 3 : (stages % 2 == 0 ? 2 : 1);
  if (const char* e = getenv("FIRE_B200_KNN_ISSUERS")) {            // A/B experiments only
    const int j = atoi(e);
    if (j >= 1 && j <= KNN_ISSUERS && stages % j == 0) p.n_issuers = j;
  }
  p.cand_score = h->cand_score; p.cand_idx = h->cand_idx;
  const size_t smem_bytes = 1024 + a_bytes + static_cast<size_t>(stages) * b_stage + bar_bytes;

  FIRE_CUDA(cudaMemsetAsync(h->counters, 0, sizeof(uint32_t) * 4, st));
  if (pair) rc = KP == 16 ? knn_launch_scan<16, true>(h, tq, tg, p, smem_bytes, st) : knn_launch_scan<64, true>(h, tq, tg, p, smem_bytes, st);
  else rc = KP == 16 ? knn_launch_scan<16, false>(h, tq, tg, p, smem_bytes, st) : knn_launch_scan<64, false>(h, tq, tg, p, smem_bytes, st);
  if (rc != FIRE_OK) return rc;

  {
    const int blocks = (Q + 7) / 8;
    if (KP == 16)
      knn_rerank_kernel<16><<<blocks, 256, 0, st>>>(h->qn32, h->g32, h->cand_score, h->cand_idx, S, D, Q, k, h->eps, h->q_err, h->g_max_err,
                                                    out, h->flagged, h->counters, h->stats);
    else
      knn_rerank_kernel<64><<<blocks, 256, 0, st>>>(h->qn32, h->g32, h->cand_score, h->cand_idx, S, D, Q, k, h->eps, h->q_err, h->g_max_err,
                                                    out, h->flagged, h->counters, h->stats);
    FIRE_LAUNCH_CHECK("knn_rerank_kernel");
    count_launch();
  }
  h->last_S = S; h->last_KP = KP; h->last_k = k; h->last_Q = Q; h->last_rows_per_split = p.rows_per_split; h->last_n_rows = n_rows;
  if (defer_fallback) return FIRE_OK;
  return knn_launch_fallback(h, out, st);
}

extern "C" {

int fire_knn_create(int D, size_t capacity, fire_knn_t** out) {
  if (!out) return fail(FIRE_ERR_ARG, "fire_knn_create: out is NULL");
  if (D <= 0 || D % 64 != 0 || D > 512) return fail(FIRE_ERR_UNSUPPORTED, "fire_knn_create: D=%d must be a multiple of 64 in [64,512]", D);
  if (capacity == 0 || capacity > 0x7FFFFFF0ull) return fail(FIRE_ERR_ARG, "fire_knn_create: capacity %zu out of range", capacity);
  fire_knn* h = new (std::nothrow) fire_knn();
  if (!h) return fail(FIRE_ERR_STATE, "out of host memory");
  h->D = D;
  h->capacity = capacity;
  if (cudaGetDevice(&h->device) != cudaSuccess || h->device < 0 || h->device >= FIRE_MAX_DEVICES) {
    delete h;
    return fail(FIRE_ERR_CUDA, "no current CUDA device (no CPU fallback)");
  }
  cudaError_t e1 = cudaMalloc(&h->g32, sizeof(float) * capacity * D);
  cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&h->g16, sizeof(__half) * capacity * D) : e1;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    cudaFree(h->g32); cudaFree(h->g16);
    delete h;
    return fail(FIRE_ERR_CUDA, "fire_knn_create: cudaMalloc of %zu x %d gallery failed: %s", capacity, D,
                cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  }
  if (cudaMalloc(&h->g_max_err, sizeof(int)) != cudaSuccess || cudaMemset(h->g_max_err, 0, sizeof(int)) != cudaSuccess) {
    cudaFree(h->g32); cudaFree(h->g16); cudaFree(h->g_max_err);
    delete h;
    return fail(FIRE_ERR_CUDA, "fire_knn_create: cudaMalloc failed");
  }
  *out = h;
  return FIRE_OK;
}

int fire_knn_destroy(fire_knn_t* h) {
  if (!h) return FIRE_OK;
  use_device(h->device);
  cudaFree(h->g32); cudaFree(h->g16); cudaFree(h->g_max_err); cudaFree(h->q_err); cudaFree(h->qn32); cudaFree(h->q16); cudaFree(h->cand_score);
  cudaFree(h->cand_idx); cudaFree(h->flagged); cudaFree(h->counters); cudaFree(h->stats);
  cudaFree(h->part_dist); cudaFree(h->part_idx); cudaFree(h->ref_dist); cudaFree(h->ref_idx); cudaFree(h->work);
  if (h->stage_q) cudaFreeHost(h->stage_q);
  if (h->stage_d) cudaFreeHost(h->stage_d);
  if (h->stage_i) cudaFreeHost(h->stage_i);
  if (h->stage_rows) cudaFree(h->stage_rows);
  if (h->host_flagged) cudaFreeHost(h->host_flagged);
  cudaFree(h->dev_q); cudaFree(h->dev_d); cudaFree(h->dev_i);
  delete h;
  return FIRE_OK;
}

int fire_knn_reset(fire_knn_t* h) {
  if (!h) return fail(FIRE_ERR_ARG, "NULL handle");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  h->count = 0;
  FIRE_CUDA(cudaMemset(h->g_max_err, 0, sizeof(int)));
  return FIRE_OK;
}
size_t fire_knn_count(const fire_knn_t* h) { return h ? h->count : 0; }
size_t fire_knn_capacity(const fire_knn_t* h) { return h ? h->capacity : 0; }
int fire_knn_dim(const fire_knn_t* h) { return h ? h->D : 0; }

int fire_knn_add(fire_knn_t* h, const float* rows, size_t n, fire_stream_t stream) {
  if (!h || (!rows && n)) return fail(FIRE_ERR_ARG, "fire_knn_add: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  if (n == 0) return FIRE_OK;
  if (h->count + n > h->capacity)
    return fail(FIRE_ERR_STATE, "fire_knn_add: %zu + %zu rows exceed capacity %zu", h->count, n, h->capacity);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t warps_per_block = 8;
  const size_t blocks = (n + warps_per_block - 1) / warps_per_block;
  knn_normalize_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(rows, n, h->D, h->g32 + h->count * h->D,
                                                                       h->g16 + h->count * h->D, n, nullptr, h->g_max_err, 0);
  FIRE_LAUNCH_CHECK("knn_normalize_kernel(add)");
  count_launch();
  h->count += n;
  return FIRE_OK;
}

int fire_knn_add_host(fire_knn_t* h, const float* host_rows, size_t n) {
  if (!h || (!host_rows && n)) return fail(FIRE_ERR_ARG, "fire_knn_add_host: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  if (n == 0) return FIRE_OK;
  if (h->count + n > h->capacity)
    return fail(FIRE_ERR_STATE, "fire_knn_add_host: %zu + %zu rows exceed capacity %zu", h->count, n, h->capacity);
  const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / (sizeof(float) * h->D));
  if (h->stage_rows_cap < std::min(n, chunk_rows)) {         // the row staging buffer is independent of the search buffers
    if (h->stage_rows) cudaFree(h->stage_rows);
    h->stage_rows = nullptr;
    h->stage_rows_cap = 0;
    FIRE_CUDA(cudaMalloc(&h->stage_rows, sizeof(float) * std::min(n, chunk_rows) * h->D));
    h->stage_rows_cap = std::min(n, chunk_rows);
  }
  for (size_t off = 0; off < n; off += h->stage_rows_cap) {
    const size_t m = std::min(h->stage_rows_cap, n - off);
    FIRE_CUDA(cudaMemcpy(h->stage_rows, host_rows + off * h->D, sizeof(float) * m * h->D, cudaMemcpyHostToDevice));
    int rc = fire_knn_add(h, h->stage_rows, m, nullptr);
    if (rc != FIRE_OK) return rc;
    FIRE_CUDA(cudaStreamSynchronize(nullptr));
  }
  return FIRE_OK;
}

int fire_knn_get_rows_host(fire_knn_t* h, size_t first, size_t n, float* host_out) {
  if (!h || (!host_out && n)) return fail(FIRE_ERR_ARG, "fire_knn_get_rows_host: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  if (first + n > h->count) return fail(FIRE_ERR_ARG, "fire_knn_get_rows_host: rows [%zu,%zu) beyond count %zu", first, first + n, h->count);
  if (n == 0) return FIRE_OK;
  FIRE_CUDA(cudaMemcpy(host_out, h->g32 + first * h->D, sizeof(float) * n * h->D, cudaMemcpyDeviceToHost));
  return FIRE_OK;
}

int fire_knn_set_margin(fire_knn_t* h, float eps) {
  if (!h) return fail(FIRE_ERR_ARG, "NULL handle");
  h->eps = eps > 0.f ? eps : KNN_DEFAULT_EPS;
  return FIRE_OK;
}

int fire_knn_stats_ex(fire_knn_t* h, uint64_t* host_out4) {
  if (!h || !host_out4) return fail(FIRE_ERR_ARG, "fire_knn_stats_ex: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  unsigned long long v[4] = {0, 0, 0, 0};
  if (h->stats) FIRE_CUDA(cudaMemcpy(v, h->stats, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) host_out4[i] = v[i];
  return FIRE_OK;
}

int fire_knn_stats(fire_knn_t* h, uint64_t* host_queries_total, uint64_t* host_queries_fallback) {
  uint64_t v[4];
  const int rc = fire_knn_stats_ex(h, v);
  if (rc != FIRE_OK) return rc;
  if (host_queries_total) *host_queries_total = v[0];
  if (host_queries_fallback) *host_queries_fallback = v[1];
  return FIRE_OK;
}

int fire_knn_search(fire_knn_t* h, const float* queries, int Q, int k, int64_t id_offset, float* out_dist,
                    int64_t* out_ids, fire_stream_t stream) {
  if (!h || !queries || !out_dist || !out_ids) return fail(FIRE_ERR_ARG, "fire_knn_search: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  KnnOut out{out_dist, reinterpret_cast<long long*>(out_ids), nullptr, id_offset, 1};
  return knn_search_impl(h, queries, -1, Q, k, out, false, static_cast<cudaStream_t>(stream));
}

int fire_knn_search_rows(fire_knn_t* h, size_t first, int Q, int k, int64_t id_offset, float* out_dist, int64_t* out_ids,
                         fire_stream_t stream) {
  if (!h || !out_dist || !out_ids) return fail(FIRE_ERR_ARG, "fire_knn_search_rows: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  if (Q <= 0 || first + static_cast<size_t>(Q) > h->count)
    return fail(FIRE_ERR_ARG, "fire_knn_search_rows: rows [%zu,%zu) beyond count %zu", first, first + static_cast<size_t>(std::max(Q, 0)), h->count);
  KnnOut out{out_dist, reinterpret_cast<long long*>(out_ids), nullptr, id_offset, 1};
  return knn_search_impl(h, nullptr, static_cast<long long>(first), Q, k, out, false, static_cast<cudaStream_t>(stream));
}

int fire_knn_search_packed(fire_knn_t* h, const float* queries, int Q, int k, int64_t id_offset, int64_t id_stride,
                           void* out_packed, fire_stream_t stream) {
  if (!h || !queries || !out_packed) return fail(FIRE_ERR_ARG, "fire_knn_search_packed: NULL argument");
  if (id_stride < 1) return fail(FIRE_ERR_ARG, "fire_knn_search_packed: id_stride=%lld", static_cast<long long>(id_stride));
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  KnnOut out{nullptr, nullptr, static_cast<int32_t*>(out_packed), id_offset, id_stride};
  return knn_search_impl(h, queries, -1, Q, k, out, true, static_cast<cudaStream_t>(stream));
}

int fire_knn_search_host(fire_knn_t* h, const float* host_queries, int Q, int k, int64_t id_offset,
                         float* host_out_dist, int64_t* host_out_ids) {
  if (!h || !host_queries || !host_out_dist || !host_out_ids) return fail(FIRE_ERR_ARG, "fire_knn_search_host: NULL argument");
  { const int rc_dev = use_device(h->device); if (rc_dev != FIRE_OK) return rc_dev; }
  if (Q <= 0 || k < 1) return fail(FIRE_ERR_ARG, "fire_knn_search_host: Q=%d k=%d", Q, k);
  const size_t qn = static_cast<size_t>(Q) * h->D, on = static_cast<size_t>(Q) * k;
  if (qn > h->stage_q_cap) {
    if (h->stage_q) cudaFreeHost(h->stage_q);
    h->stage_q = nullptr; h->stage_q_cap = 0;
    FIRE_CUDA(cudaMallocHost(&h->stage_q, sizeof(float) * qn));
    h->stage_q_cap = qn;
  }
  if (on > h->stage_o_cap) {
    if (h->stage_d) cudaFreeHost(h->stage_d);
    if (h->stage_i) cudaFreeHost(h->stage_i);
    h->stage_d = nullptr; h->stage_i = nullptr; h->stage_o_cap = 0;
    FIRE_CUDA(cudaMallocHost(&h->stage_d, sizeof(float) * on));
    FIRE_CUDA(cudaMallocHost(&h->stage_i, sizeof(long long) * on));
    h->stage_o_cap = on;
  }
  if (qn > h->dev_q_cap) {
    cudaFree(h->dev_q); h->dev_q = nullptr; h->dev_q_cap = 0;
    FIRE_CUDA(cudaMalloc(&h->dev_q, sizeof(float) * qn));
    h->dev_q_cap = qn;
  }
  if (on > h->dev_o_cap) {
    cudaFree(h->dev_d); cudaFree(h->dev_i); h->dev_d = nullptr; h->dev_i = nullptr; h->dev_o_cap = 0;
    FIRE_CUDA(cudaMalloc(&h->dev_d, sizeof(float) * on));
    FIRE_CUDA(cudaMalloc(&h->dev_i, sizeof(long long) * on));
    h->dev_o_cap = on;
  }
  if (!h->host_flagged) FIRE_CUDA(cudaMallocHost(&h->host_flagged, sizeof(uint32_t)));
  memcpy(h->stage_q, host_queries, sizeof(float) * qn);
  FIRE_CUDA(cudaMemcpyAsync(h->dev_q, h->stage_q, sizeof(float) * qn, cudaMemcpyHostToDevice, nullptr));
  // This call returns host arrays, so it synchronises anyway: the flagged count comes back with the results and the four
  // fallback kernels are launched only when a query's proof failed (rare) - they cost a single query ~15 us of launch latency.
  KnnOut out{h->dev_d, h->dev_i, nullptr, id_offset, 1};
  *h->host_flagged = 0;
  int rc = knn_search_impl(h, h->dev_q, -1, Q, k, out, false, nullptr, true);
  if (rc != FIRE_OK) return rc;
  FIRE_CUDA(cudaMemcpyAsync(h->stage_d, h->dev_d, sizeof(float) * on, cudaMemcpyDeviceToHost, nullptr));
  FIRE_CUDA(cudaMemcpyAsync(h->stage_i, h->dev_i, sizeof(long long) * on, cudaMemcpyDeviceToHost, nullptr));
  if (h->last_Q > 0) FIRE_CUDA(cudaMemcpyAsync(h->host_flagged, h->counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, nullptr));
  FIRE_CUDA(cudaStreamSynchronize(nullptr));
  if (h->last_Q > 0 && *h->host_flagged > 0) {
    rc = knn_launch_fallback(h, out, nullptr);
    if (rc != FIRE_OK) return rc;
    FIRE_CUDA(cudaMemcpyAsync(h->stage_d, h->dev_d, sizeof(float) * on, cudaMemcpyDeviceToHost, nullptr));
    FIRE_CUDA(cudaMemcpyAsync(h->stage_i, h->dev_i, sizeof(long long) * on, cudaMemcpyDeviceToHost, nullptr));
    FIRE_CUDA(cudaStreamSynchronize(nullptr));
  }
  memcpy(host_out_dist, h->stage_d, sizeof(float) * on);
  memcpy(host_out_ids, h->stage_i, sizeof(long long) * on);
  return FIRE_OK;
}

static int knn_merge_impl(const float* dists, const int64_t* ids, const void* packed, int Q, int k, int G, float* out_dist, int64_t* out_ids,
                          fire_stream_t stream) {
  if (!out_dist || !out_ids) return fail(FIRE_ERR_ARG, "fire_knn_merge: NULL argument");
  if (Q <= 0 || k < 1 || G < 1 || G > 32 * KNN_MAX_LISTS_PER_LANE)
    return fail(FIRE_ERR_ARG, "fire_knn_merge: Q=%d k=%d G=%d out of range", Q, k, G);
  const int blocks = (Q + 7) / 8;
  knn_merge_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(dists, reinterpret_cast<const long long*>(ids),
                                                                           static_cast<const int32_t*>(packed), Q, k, G, out_dist,
                                                                           reinterpret_cast<long long*>(out_ids));
  FIRE_LAUNCH_CHECK("knn_merge_kernel");
  count_launch();
  return FIRE_OK;
}

int fire_knn_merge(const float* dists, const int64_t* ids, int Q, int k, int G, float* out_dist, int64_t* out_ids,
                   fire_stream_t stream) {
  if (!dists || !ids) return fail(FIRE_ERR_ARG, "fire_knn_merge: NULL argument");
  return knn_merge_impl(dists, ids, nullptr, Q, k, G, out_dist, out_ids, stream);
}

int fire_knn_merge_packed(const void* packed, int Q, int k, int G, float* out_dist, int64_t* out_ids, fire_stream_t stream) {
  if (!packed) return fail(FIRE_ERR_ARG, "fire_knn_merge_packed: NULL argument");
  return knn_merge_impl(nullptr, nullptr, packed, Q, k, G, out_dist, out_ids, stream);
}

}  // extern "C"
