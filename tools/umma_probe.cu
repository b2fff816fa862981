// umma_probe.cu - hardware probes that decide the conv-stack design (run on a B200 through gpurun).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
//
// Part 1  L2 -> SM delivery bandwidth: every SM streams an L2-resident region (LDG.128 and cp.async.bulk),
//         distinct slices per CTA and the same slice for all CTAs (the "weights re-streamed per tile" pattern).
// Part 2  tcgen05.mma with a SHIFTED A operand: the A rows are consecutive 16-byte (or 32/64/128-byte) records of a
//         longer "position" array in shared memory and the descriptor start address is advanced by s records, which
//         is what a stride-1 k x k convolution needs to read tap (r, s) without materialising im2col.  Variants:
//         no-swizzle planar-8 layout (LBO/SBO both ways), SWIZZLE_32B/64B/128B with and without base_offset.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../face-identification-in-real-time-environments-fire_b200/csrc/fire_common.cuh"

using namespace fire;

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)

// ------------------------------------------------------------------------------------------------ part 1
__global__ void __launch_bounds__(512, 1) l2_ldg_kernel(const uint4* __restrict__ buf, size_t n16_per_cta, int same, int iters,
                                                        unsigned* sink) {
  const uint4* p = buf + (same ? 0 : static_cast<size_t>(blockIdx.x) * n16_per_cta);
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    for (size_t i = threadIdx.x; i + 7 * 512 < n16_per_cta; i += 8 * 512) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(p + i + j * 512);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one producer thread keeps `depth` bulk copies of `chunk` bytes in flight
__global__ void __launch_bounds__(128, 1) l2_bulk_kernel(const uint8_t* __restrict__ buf, size_t bytes_per_cta, int same, int iters,
                                                         int chunk, int depth) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* p = buf + (same ? 0 : static_cast<size_t>(blockIdx.x) * bytes_per_cta);
    const long long n = static_cast<long long>(bytes_per_cta / chunk) * iters;
    const long long per_iter = bytes_per_cta / chunk;
    for (long long i = 0; i < n + depth; ++i) {
      const int s = static_cast<int>(i % depth);
      if (i >= depth) mbar_wait(&bars[s], static_cast<uint32_t>(((i - depth) / depth) & 1), 1);
      if (i < n) {
        mbar_arrive_expect_tx(&bars[s], chunk);
        bulk_g2s(smem_u32(smem + static_cast<size_t>(s) * chunk), p + (i % per_iter) * chunk, chunk, &bars[s]);
      }
    }
  }
}

static void part1() {
  printf("== part 1: L2 -> SM delivery bandwidth (148 CTAs) ==\n");
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t total = 64ull << 20;                 // 64 MB: L2 resident
  uint8_t* buf;
  unsigned* sink;
  CK(cudaMalloc(&buf, 2048ull << 20));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, 2048ull << 20));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  CK(cudaFuncSetAttribute(l2_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Case { const char* name; size_t region; int same; };
  const Case cases[] = {{"distinct slices, 64 MB region (L2 resident)", total, 0},
                        {"same 256 KB slice for every CTA (L2 broadcast)", 256 << 10, 1},
                        {"distinct slices, 2 GB region (HBM)", 2048ull << 20, 0}};
  for (const Case& c : cases) {
    const size_t per_cta = c.same ? c.region : (c.region / sms) & ~size_t(65535);
    const int iters = c.same ? 256 : (c.region > (512ull << 20) ? 1 : 16);
    const double bytes = static_cast<double>(per_cta) * sms * iters;
    float ms;
    // LDG
    l2_ldg_kernel<<<sms, 512>>>(reinterpret_cast<const uint4*>(buf), per_cta / 16, c.same, 1, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    l2_ldg_kernel<<<sms, 512>>>(reinterpret_cast<const uint4*>(buf), per_cta / 16, c.same, iters, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  LDG.128 x8 unroll, 512 thr : %-50s %8.1f GB/s  (%.1f B/clk/SM at 1.965 GHz)\n", c.name, bytes / ms / 1e6,
           bytes / ms / 1e6 / sms / 1.965);
    for (int chunk : {4096, 16384}) {
      for (int depth : {4, 8}) {
        if (static_cast<size_t>(chunk) * depth > 190 * 1024) continue;
        l2_bulk_kernel<<<sms, 128, static_cast<size_t>(chunk) * depth + 256>>>(buf, per_cta, c.same, 1, chunk, depth);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        l2_bulk_kernel<<<sms, 128, static_cast<size_t>(chunk) * depth + 256>>>(buf, per_cta, c.same, iters, chunk, depth);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        printf("  bulk copy %5d B x depth %d   : %-50s %8.1f GB/s  (%.1f B/clk/SM)\n", chunk, depth, c.name, bytes / ms / 1e6,
               bytes / ms / 1e6 / sms / 1.965);
      }
    }
  }
  cudaFree(buf);
  cudaFree(sink);
}

// ------------------------------------------------------------------------------------------------ part 2
constexpr int NPOS = 512;      // positions (records) in the smem array
constexpr int PN = 32;         // UMMA N
struct ProbeCfg {
  int layout;        // 0 = no swizzle planar-8, 1 = SW32, 2 = SW64, 3 = SW128
  int shift;         // A rows = positions shift .. shift+127
  int swap_lbo_sbo;  // no-swizzle only
  int base_off_mode; // 0: field = 0; 1: field = (start_addr >> 7) & 7
  int C;             // channels per position (K of the product)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_code, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(layout_code) << 61;
  return d;
}

// pos[NPOS][C] fp16 (global, position-major), w[PN][C] fp16 -> out[128][PN] fp32 = sum_c pos[m + shift][c] * w[n][c]
__global__ void __launch_bounds__(128, 1) umma_shift_kernel(const __half* __restrict__ pos, const __half* __restrict__ w, float* out,
                                                            ProbeCfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // NPOS * C * 2 bytes (<= 64 KB)
  uint8_t* sB = smem + 64 * 1024;                       // PN rows x 128 B, SW128, K padded to 64
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int C = cfg.C, chunks = C / 8;
  const int rowbytes = C * 2;
  // ---- fill A
  for (int i = threadIdx.x; i < NPOS * chunks; i += blockDim.x) {
    const int p = i / chunks, c = i % chunks;
    const uint4 v = *reinterpret_cast<const uint4*>(pos + static_cast<size_t>(p) * C + c * 8);
    uint32_t off;
    if (cfg.layout == 0) off = static_cast<uint32_t>(c * (NPOS * 16) + p * 16);                    // planar-8
    else if (cfg.layout == 1) off = static_cast<uint32_t>(p * 32 + ((c ^ ((p >> 2) & 1)) << 4));   // SW32: C = 16
    else if (cfg.layout == 2) off = static_cast<uint32_t>(p * 64 + ((c ^ ((p >> 1) & 3)) << 4));   // SW64: C = 32
    else off = static_cast<uint32_t>(p * 128 + ((c ^ (p & 7)) << 4));                              // SW128: C = 64
    *reinterpret_cast<uint4*>(sA + off) = v;
  }
  // ---- fill B (zero padded to K = 64)
  for (int i = threadIdx.x; i < PN * 8; i += blockDim.x) {
    const int n = i / 8, c = i % 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c < chunks) v = *reinterpret_cast<const uint4*>(w + static_cast<size_t>(n) * C + c * 8);
    *reinterpret_cast<uint4*>(sB + sw128_offset(n, c)) = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<32>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, PN);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    for (int ks = 0; ks < C / 16; ++ks) {
      uint64_t ad;
      if (cfg.layout == 0) {
        const uint32_t start = a_base + static_cast<uint32_t>(ks * 2 * (NPOS * 16) + cfg.shift * 16);
        const uint32_t k_stride = NPOS * 16, m_stride = 128;
        ad = cfg.swap_lbo_sbo ? make_desc(start, m_stride, k_stride, 0, 0) : make_desc(start, k_stride, m_stride, 0, 0);
      } else {
        const uint32_t start = a_base + static_cast<uint32_t>(cfg.shift * rowbytes + ks * 32);
        const uint32_t code = cfg.layout == 1 ? 6u : cfg.layout == 2 ? 4u : 2u;
        const uint32_t bo = cfg.base_off_mode ? ((start >> 7) & 7) : 0u;
        ad = make_desc(start, 16, static_cast<uint32_t>(8 * rowbytes), code, bo);
      }
      umma_f16(tmem, ad, umma_desc_sw128(b_base + ks * 32), idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 2);
  tc_fence_after();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_ld_wait(r);
    for (int j = 0; j < PN; ++j) out[(warp * 32 + lane) * PN + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<32>(tmem); }
}

static void part2() {
  printf("== part 2: tcgen05.mma with a row-shifted A descriptor ==\n");
  CK(cudaFuncSetAttribute(umma_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024 + 1024));
  const char* lname[] = {"none/planar-8", "SW32", "SW64", "SW128"};
  const int layC[] = {32, 16, 32, 64};
  for (int layout = 0; layout < 4; ++layout) {
    const int C = layC[layout];
    std::vector<__half> hpos(NPOS * C), hw(PN * C);
    srand(1234 + layout);
    for (auto& v : hpos) v = __float2half(static_cast<float>(rand() % 7 - 3));
    for (auto& v : hw) v = __float2half(static_cast<float>(rand() % 5 - 2));
    __half *dpos, *dw;
    float* dout;
    CK(cudaMalloc(&dpos, hpos.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&dout, 128 * PN * 4));
    CK(cudaMemcpy(dpos, hpos.data(), hpos.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    const int n_var = layout == 0 ? 1 : 2;   // the swapped LBO/SBO assignment reads out of bounds (measured: illegal address)
    for (int var = 0; var < n_var; ++var) {
      printf("  layout %-14s %s:", lname[layout],
             layout == 0 ? (var ? "LBO=M-stride SBO=K-stride" : "LBO=K-stride SBO=M-stride") : (var ? "base_offset=(addr>>7)&7" : "base_offset=0          "));
      for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 9, 19, 79, 160, 161}) {
        ProbeCfg cfg{layout, shift, layout == 0 ? var : 0, layout == 0 ? 0 : var, C};
        CK(cudaMemset(dout, 0xff, 128 * PN * 4));
        umma_shift_kernel<<<1, 128, 72 * 1024 + 1024>>>(dpos, dw, dout, cfg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" [shift %d: CUDA error %s]\n", shift, cudaGetErrorString(e)); exit(3); }
        std::vector<float> hout(128 * PN);
        CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < PN; ++n) {
            float ref = 0.f;
            for (int c = 0; c < C; ++c) ref += __half2float(hpos[(m + shift) * C + c]) * __half2float(hw[n * C + c]);
            if (ref != hout[m * PN + n]) ++bad;
          }
        printf(" s%d:%s", shift, bad ? "FAIL" : "ok");
        if (bad) printf("(%d)", bad);
      }
      printf("\n");
    }
    cudaFree(dpos); cudaFree(dw); cudaFree(dout);
  }
}


// ------------------------------------------------------------------------------------------------ part 3
// How fast can ONE SM pull L2-resident bytes into shared memory?  (a) 2-D tensor TMA, box 64 x R fp16 rows (SW128),
// `nthr` issuing threads (one per warp) each with `depth` loads in flight; (b) cp.async 16 B by `nthr*32` threads.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(256, 1) tma2d_kernel(const __grid_constant__ CUtensorMap tmap, int rows_per_cta, int box_rows,
                                                       int depth, int nthr, int iters, int lane_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * nthr; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = lane_mode ? static_cast<int>(threadIdx.x) : static_cast<int>(threadIdx.x >> 5);   // issuer index
  if ((lane_mode ? threadIdx.x < nthr : ((threadIdx.x & 31) == 0 && warp < nthr))) {
    const uint32_t bytes = box_rows * 128;
    const int loads_per_iter = rows_per_cta / box_rows / nthr;       // this thread's share
    const long long n = static_cast<long long>(loads_per_iter) * iters;
    uint64_t* mybar = bars + warp * depth;
    uint8_t* mysm = smem + static_cast<size_t>(warp) * depth * bytes;
    const int row0 = blockIdx.x * rows_per_cta + warp * loads_per_iter * box_rows;
    for (long long i = 0; i < n + depth; ++i) {
      const int s = static_cast<int>(i % depth);
      if (i >= depth) mbar_wait(&mybar[s], static_cast<uint32_t>(((i - depth) / depth) & 1), 3);
      if (i < n) {
        mbar_arrive_expect_tx(&mybar[s], bytes);
        tma_load_2d(mysm + static_cast<size_t>(s) * bytes, &tmap, &mybar[s], 0, row0 + static_cast<int>(i % loads_per_iter) * box_rows);
      }
    }
  }
}

__global__ void __launch_bounds__(512, 1) cpasync_kernel(const uint8_t* __restrict__ buf, size_t bytes_per_cta, int iters, int groups_in_flight) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint8_t* p = buf + static_cast<size_t>(blockIdx.x) * bytes_per_cta;
  // each "group" = 4 x 16 B per thread = 64 B x blockDim; ring of 8 group slots in smem
  const size_t group_bytes = static_cast<size_t>(blockDim.x) * 64;
  const long long n = static_cast<long long>(bytes_per_cta / group_bytes) * iters;
  const long long per_iter = bytes_per_cta / group_bytes;
  for (long long i = 0; i < n; ++i) {
    const int s = static_cast<int>(i & 7);
    const uint8_t* src = p + (i % per_iter) * group_bytes;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cp_async_16(smem_u32(smem + s * group_bytes + (j * blockDim.x + threadIdx.x) * 16), src + (j * blockDim.x + threadIdx.x) * 16, true);
    cp_async_commit();
    cp_async_wait_dyn(groups_in_flight - 1);
  }
  cp_async_wait_all();
}

static void part3() {
  printf("== part 3: per-SM fill rate of shared memory from an L2-resident 64 MB region (148 CTAs) ==\n");
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  const size_t total = 64ull << 20;
  uint8_t* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  const uint64_t rows = total / 128;                      // [rows, 64] fp16, dense
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  CK(cudaFuncSetAttribute(tma2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(cpasync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int rows_per_cta = static_cast<int>((rows / sms) / 1024 * 1024);
  const int iters = 16;
  const double bytes = static_cast<double>(rows_per_cta) * 128 * sms * iters;
  for (int box_rows : {32, 128, 256}) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {64, rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    for (int lane_mode : {0, 1})
    for (int nthr : {1, 2, 4, 8}) {
      if (lane_mode && nthr == 1) continue;
      for (int depth : {2, 4}) {
        if (depth * nthr > 64) continue;
        const size_t sm = static_cast<size_t>(box_rows) * 128 * depth * nthr + 1024;
        if (sm > 196 * 1024) continue;
        tma2d_kernel<<<sms, 256, sm>>>(tm, rows_per_cta, box_rows, depth, nthr, 1, lane_mode);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tma2d_kernel<<<sms, 256, sm>>>(tm, rows_per_cta, box_rows, depth, nthr, iters, lane_mode);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("  TMA 2D box 64x%-3d (%5d B) issuers %d (%s) depth %d : %8.1f GB/s  (%.1f B/clk/SM)\n", box_rows, box_rows * 128, nthr, lane_mode ? "lanes of one warp" : "one per warp", depth,
               bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.965);
      }
    }
  }
  for (int threads : {128, 256, 384}) {
    for (int gif : {2, 4, 8}) {
      const size_t per_cta = static_cast<size_t>(rows_per_cta) * 128;
      cpasync_kernel<<<sms, threads, 8 * threads * 64 + 1024>>>(buf, per_cta, 1, gif);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      cpasync_kernel<<<sms, threads, 8 * threads * 64 + 1024>>>(buf, per_cta, iters, gif);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("  cp.async 16 B x4/thread, %3d threads, %d groups in flight : %8.1f GB/s  (%.1f B/clk/SM)\n", threads, gif, bytes / ms / 1e6,
             bytes / ms / 1e6 / sms / 1.965);
    }
  }
  cudaFree(buf);
}

// ------------------------------------------------------------------------------------------------ part 4
// Are TMA loads from ONE thread pipelined?  Issue n loads back to back (no waits in between), then wait for all:
// t(n) ~ t(1) means pipelined, t(n) ~ n * t(1) means the hardware runs them one at a time.  wait_mode: 0 = try_wait with
// the suspend-time hint used by fire_common.cuh, 1 = plain test_wait spin.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(128, 1) tma_pipe_kernel(const __grid_constant__ CUtensorMap tmap, int box_rows, int n, int wait_mode,
                                                          int one_barrier, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = box_rows * 128;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 6; ++rep) {
      const uint32_t par = rep & 1;
      const long long t0 = clock64();
      if (one_barrier) {
        mbar_arrive_expect_tx(&bars[0], bytes * n);
        for (int i = 0; i < n; ++i)
          tma_load_2d(smem + static_cast<size_t>(i) * bytes, &tmap, &bars[0], 0, (blockIdx.x * 16 + i) * box_rows);
        if (wait_mode == 0) mbar_wait(&bars[0], par, 5); else while (!mbar_test_wait(&bars[0], par)) {}
      } else {
        for (int i = 0; i < n; ++i) {
          mbar_arrive_expect_tx(&bars[i], bytes);
          tma_load_2d(smem + static_cast<size_t>(i) * bytes, &tmap, &bars[i], 0, (blockIdx.x * 16 + i) * box_rows);
        }
        for (int i = 0; i < n; ++i) {
          if (wait_mode == 0) mbar_wait(&bars[i], par, 5); else while (!mbar_test_wait(&bars[i], par)) {}
        }
      }
      const long long t1 = clock64();
      if (rep >= 2 && t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
  }
}

static void part4() {
  printf("== part 4: n back-to-back TMA loads from one thread, cycles until all have landed (min of 4 reps, L2 hits) ==\n");
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  const size_t total = 256ull << 20;
  uint8_t* buf;
  long long* dout;
  CK(cudaMalloc(&buf, total));
  CK(cudaMalloc(&dout, 148 * 8));
  CK(cudaMemset(buf, 1, total));
  CK(cudaFuncSetAttribute(tma_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int box_rows : {32, 128}) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {64, total / 128};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    for (int grid : {1, 148})
      for (int wait_mode : {0, 1})
        for (int one_barrier : {0, 1}) {
          printf("  box 64x%-3d grid %3d %s %s :", box_rows, grid, wait_mode ? "test_wait spin " : "try_wait+suspend", one_barrier ? "one barrier " : "barrier/load");
          for (int n : {1, 2, 4, 8}) {
            tma_pipe_kernel<<<grid, 128, 8 * box_rows * 128 + 1024>>>(tm, box_rows, n, wait_mode, one_barrier, dout);
            CK(cudaDeviceSynchronize());
            long long h[148];
            CK(cudaMemcpy(h, dout, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("  n=%d: %5lld", n, mx);
          }
          printf("\n");
        }
  }
  cudaFree(buf);
  cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ part 5
// Cost of a chain of tcgen05.mma (M=128, N, K=16) into ONE accumulator as a function of the A layout and of the
// row shift of the A descriptor: cycles from the first issue until the commit barrier fires.
__global__ void __launch_bounds__(128, 1) umma_time_kernel(int layout, int N, int n_mma, int shift_rows, int wbox, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 64 KB of zeros
  uint8_t* sB = smem + 64 * 1024;     // 64 KB of zeros
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t rowbytes = layout == 3 ? 128u : layout == 2 ? 64u : 32u;
    const uint32_t code = layout == 3 ? 2u : layout == 2 ? 4u : 6u;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      const long long t0 = clock64();
      for (int j = 0; j < n_mma; ++j) {
        // tap-like walk: shift = shift_rows * ((j / 2) % 3) + wbox * ((j / 2) / 3), k half = j & 1
        const int tap = j >> 1;
        const uint32_t rows = static_cast<uint32_t>(shift_rows * (tap % 3) + wbox * (tap / 3));
        const uint32_t a_addr = smem_u32(sA) + rows * rowbytes + (j & 1) * 32;
        const uint64_t ad = layout == 0 ? umma_desc_nosw(smem_u32(sA) + rows * 16, 8192, 128) : make_desc(a_addr, 16, 8 * rowbytes, code, 0);
        umma_f16(tmem, ad, umma_desc_sw128(smem_u32(sB) + (j & 3) * 32), idesc, j ? 1u : 0u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1, 7);
      const long long t1 = clock64();
      if (rep >= 1 && t1 - t0 < best) best = t1 - t0;
    }
    out[0] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

static void part5() {
  printf("== part 5: cycles for a chain of n tcgen05.mma (M=128, K=16) + commit, by A layout / row shift ==\n");
  CK(cudaFuncSetAttribute(umma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024));
  long long* dout;
  CK(cudaMalloc(&dout, 8));
  const char* lname[] = {"none/planar-8", "SW32", "SW64", "SW128"};
  for (int N : {32, 64, 256}) {
    for (int layout : {3, 2, 0}) {
      for (int shift : {0, 1, 8}) {
        printf("  N=%-3d A layout %-14s tap shift %d row(s), patch width %2d:", N, lname[layout], shift, shift ? 79 : 0);
        for (int n : {1, 2, 18, 36}) {
          umma_time_kernel<<<1, 128, 130 * 1024>>>(layout, N, n, shift, shift ? 79 : 0, dout);
          CK(cudaDeviceSynchronize());
          long long h;
          CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
          printf("  n=%-2d %5lld", n, h);
        }
        printf("\n");
      }
    }
  }
  cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ part 6
// Is the ~100-cycle cost of a small-N tcgen05.mma an accumulator-dependency latency or an issue limit?
// n_acc independent accumulators are fed round-robin by n_thr issuing threads (one per warp).
__global__ void __launch_bounds__(128, 1) umma_ilp_kernel(int N, int n_mma, int n_acc, int n_thr, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * 1024;
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < n_thr) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      const long long t0 = clock64();
      // this thread owns accumulators a = warp, warp + n_thr, ...; n_mma MMAs per accumulator, interleaved
      for (int j = 0; j < n_mma; ++j)
        for (int a = warp; a < n_acc; a += n_thr)
          umma_f16(tmem + a * N, umma_desc_sw128(smem_u32(sA) + (j & 3) * 32 + a * 16384), umma_desc_sw128(smem_u32(sB) + (j & 3) * 32), idesc,
                   j ? 1u : 0u);
      umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], rep & 1, 8);
      const long long t1 = clock64();
      if (rep >= 1 && t1 - t0 < best) best = t1 - t0;
    }
    out[warp] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

static void part6() {
  printf("== part 6: n tcgen05.mma (M=128, K=16) per accumulator, n_acc accumulators, n_thr issuing threads: cycles (max over threads) ==\n");
  CK(cudaFuncSetAttribute(umma_ilp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024));
  long long* dout;
  CK(cudaMalloc(&dout, 32));
  for (int N : {32, 64, 128}) {
    for (int n_acc : {1, 2, 4}) {
      for (int n_thr : {1, 2, 4}) {
        if (n_thr > n_acc || n_acc * N > 512) continue;
        printf("  N=%-3d accumulators %d issuing threads %d:", N, n_acc, n_thr);
        for (int n : {18, 36}) {
          umma_ilp_kernel<<<1, 128, 130 * 1024>>>(N, n, n_acc, n_thr, dout);
          CK(cudaDeviceSynchronize());
          long long h[4];
          CK(cudaMemcpy(h, dout, 32, cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (int i = 0; i < n_thr; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("  n=%-2d %5lld (%.0f clk per MMA)", n, mx, (double)mx / (n * n_acc));
        }
        printf("\n");
      }
    }
  }
  cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ part 7
// cta_group::2: does ONE tcgen05.mma that spans a CTA pair (M = 256, each SM holds 128 rows of A and half of B) cost the
// same ~85-130 cycles as the single-CTA M = 128 instruction?  If so, the small-N strip layers (N = 32 / 64) would do twice
// the work per instruction.  Leader CTA issues, tcgen05.commit multicasts the completion to both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_2cta_kernel(int N, int n_mma, int n_acc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      if (rank == 0) {
        const uint32_t idesc = umma_idesc_f16(256, N);
        const long long t0 = clock64();
        for (int j = 0; j < n_mma; ++j)
          for (int a = 0; a < n_acc; ++a)
            umma_f16_2cta(tmem + a * N, umma_desc_sw128(smem_u32(sA) + (j & 3) * 32 + a * 16384), umma_desc_sw128(smem_u32(sB) + (j & 3) * 32), idesc,
                          j ? 1u : 0u);
        umma_commit_2cta(&bar);
        mbar_wait(&bar, rep & 1, 9);
        const long long t1 = clock64();
        if (rep >= 1 && t1 - t0 < best) best = t1 - t0;
      } else {
        mbar_wait(&bar, rep & 1, 10);                 // the multicast commit arrives here too
      }
    }
    if (rank == 0) out[blockIdx.x / 2] = best;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static void part7() {
  printf("== part 7: n tcgen05.mma.cta_group::2 (M=256 over a CTA pair, K=16) per accumulator, n_acc accumulators, one issuing thread ==\n");
  CK(cudaFuncSetAttribute(umma_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024));
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  for (int N : {32, 64, 128, 256}) {
    for (int n_acc : {1, 2, 4}) {
      if (n_acc * N > 512) continue;
      printf("  N=%-3d accumulators %d:", N, n_acc);
      for (int n : {18, 36}) {
        umma_2cta_kernel<<<2, 128, 130 * 1024>>>(N, n, n_acc, dout);
        CK(cudaDeviceSynchronize());
        long long h = 0;
        CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
        printf("  n=%-2d %5lld (%.0f clk per MMA = per 2 x 128 rows)", n, h, (double)h / (n * n_acc));
      }
      printf("\n");
    }
  }
  cudaFree(dout);
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 3;
  CK(cudaSetDevice(0));
  if (which & 2) part2();
  if (which & 1) part1();
  if (which & 4) part3();
  if (which & 8) part4();
  if (which & 16) part5();
  if (which & 32) part6();
  if (which & 64) part7();
  return 0;
}
