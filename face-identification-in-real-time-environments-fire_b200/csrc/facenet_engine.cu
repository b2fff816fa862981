// facenet_engine.cu - executes the FaceNet plan (fire_b200/netplan.py) on one B200.
//
// Replaces onnxruntime's InferenceSession.run for weights/facenet{128,512}.onnx
// (reference facenet_gpu.py:72,116-129).  The blob produced by fire_b200.weights.pack() carries
// the op list, the buffer table (per-image offsets into one workspace arena, live ranges already
// resolved) and the BN-folded fp16 weights; this file uploads the weights once, builds their TMA
// descriptors once, and on every forward() enqueues one kernel per op on the caller's stream:
//   conv_igemm_kernel_t x 19 (single CTAs, or CTA pairs with cta_group::2 where the tile pairs fill the GPU) + conv_strip_kernel_t x 3
//   + block35_fused_kernel x 1 + block17_fused_kernel x 1 + block8_fused_kernel x 6 (tcgen05; the fused kernels run the 20 / 40 convs
//   of the five Block35 / ten Block17 blocks and the 1x3 -> 3x1 -> up tail of every Block8 block, the last one also the average pool),
//   maxpool3x3s2_kernel x 3 (the two pool branches on a side stream),  l2norm_kernel x 1.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "conv_igemm.cuh"
#include "conv_strip.cuh"
#include "block17_fused.cuh"
#include "block35_fused.cuh"
#include "block8_fused.cuh"
#include "pool_conv_fused.cuh"
#include "fire_internal.h"

namespace fire {

constexpr int OP_CONV = 1, OP_MAXPOOL = 2, OP_GAP = 3;
constexpr uint32_t BLOB_VERSION = 5;

#pragma pack(push, 1)
struct BlobHeader {
  char magic[8];
  int32_t version, D, n_ops, n_bufs;
  int64_t ws_bytes_per_image, weights_off, weights_bytes;
  int32_t in_buf, out_buf;
};
struct BlobBuf {
  int32_t H, W, C, elt;
  int64_t offset;
  int32_t external, Wp;     // external: 1 = bound by the caller, 2 = a view of the network input; Wp: row pitch in pixels (0 = W); > W for the outputs of flat-mode strip convs
};
struct BlobOp {
  int32_t kind, src_buf, src_coff, dst_buf, dst_coff, res_buf, res_coff, H, W, Ho, Wo, kh, kw, stride, pad_h, pad_w, cin,
      cout, k_pad, flags, bn_tile, flop_k;      // flop_k: real K for FLOP accounting (0 = kh * kw * cin)
  int64_t w_off, b_off;
};
#pragma pack(pop)
static_assert(sizeof(BlobHeader) == 56, "header layout must match weights.HEADER_DT");
static_assert(sizeof(BlobBuf) == 32, "buffer layout must match weights.BUF_DT");
static_assert(sizeof(BlobOp) == 104, "op layout must match weights.OP_DT");

// ---------------------------------------------------------------------------------------------
// 3x3 stride-2 VALID max-pool, NHWC fp16, 8 channels (16 bytes) per thread, channel-offset store.
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, int in_ld, int in_coff, __half* __restrict__ out,
                                    int out_ld, int out_coff, int B, int H, int W, int Ho, int Wo, int C) {   // W = input row pitch
  const int c8 = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    long long pix = i / c8;
    const int wo = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int ho = static_cast<int>(pix % Ho);
    const int b = static_cast<int>(pix / Ho);
    __half2 m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = __float2half2_rn(-65504.f);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const size_t off = (static_cast<size_t>(b * H + ho * 2 + r) * W + (wo * 2 + s)) * in_ld + in_coff + cc * 8;
        uint4 q = __ldg(reinterpret_cast<const uint4*>(in + off));
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], h[j]);
      }
    const size_t o = (static_cast<size_t>(b * Ho + ho) * Wo + wo) * out_ld + out_coff + cc * 8;
    *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<uint4*>(m);
  }
}

// global average pool over HW pixels (fp32 accumulation), NHWC fp16 -> [B, C] fp16
__global__ void gap_kernel(const __half* __restrict__ in, int in_ld, int in_coff, __half* __restrict__ out, int out_ld,
                           int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;    // one thread per (b, channel pair)
  const int c2 = C >> 1;
  if (i >= B * c2) return;
  const int b = i / c2, c = (i - b * c2) * 2;
  float s0 = 0.f, s1 = 0.f;
  for (int p = 0; p < HW; ++p) {
    float2 f = __half22float2(*reinterpret_cast<const __half2*>(in + (static_cast<size_t>(b) * HW + p) * in_ld + in_coff + c));
    s0 += f.x; s1 += f.y;
  }
  const float inv = 1.0f / static_cast<float>(HW);
  *reinterpret_cast<__half2*>(out + static_cast<size_t>(b) * out_ld + c) = __floats2half2_rn(s0 * inv, s1 * inv);
}

// rows / ||row||_2 (face_recognition.py:225-229); a zero row stays zero (the caller skips such faces)
__global__ void l2norm_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int D) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) { float v = in[static_cast<size_t>(row) * D + c]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float n = sqrtf(s);
  const float inv = n > 0.f ? 1.0f / n : 0.f;
  for (int c = lane; c < D; c += 32) {
    float v = in[static_cast<size_t>(row) * D + c];
    out[static_cast<size_t>(row) * D + c] = n > 0.f ? v * inv : v;
  }
}

// float NHWC3 in [0,1] -> the network input: fp16 space-to-depth [80][80][16], pixel scale (x*255; exact for x = k/255),
// channel (dy*2+dx)*3+c of position (Y,X) = pixel (2Y+dy, 2X+dx), channels 12..15 zero.  One thread per position.
__global__ void ingest_f32_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n_pos) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pos;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / 6400;
    const int r = static_cast<int>(i - b * 6400), Y = r / 80, X = r - Y * 80;
    float v[12];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float* src = in + ((b * 160 + 2 * Y + dy) * 160 + 2 * X) * 3;      // two adjacent pixels = 6 floats
#pragma unroll
      for (int k = 0; k < 6; ++k) v[dy * 6 + k] = src[k] * 255.0f;
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = make_uint4(pack_f16x2_sat(v[0], v[1]), pack_f16x2_sat(v[2], v[3]), pack_f16x2_sat(v[4], v[5]), pack_f16x2_sat(v[6], v[7]));
    o[1] = make_uint4(pack_f16x2_sat(v[8], v[9]), pack_f16x2_sat(v[10], v[11]), 0u, 0u);
  }
}

}  // namespace fire

using namespace fire;

struct OpRt {
  BlobOp op;
  CUtensorMap tmap_w;     // weights [cout][k_pad], box rows = bn_tile (rebuilt when the tiling changes with B)
  CUtensorMap tmap_a;     // activation matrix for TMA-mode convs (rebuilt when pointers / B change)
  CUtensorMap tmap_res;   // residual matrix (up convs): box 64 columns x 128 rows
  CUtensorMap tmap_out;   // output slice, box = box_cols x 32 rows (TMA store)
  bool tma_a = false;
  bool im2col = false;    // k x k layer whose A operand comes through an im2col tensor map (ConvParams::tma_a == 2)
  int bn_tile = 0, stages = 0, n_issuers = 1, tmem_cols = 0, m_tiles = 0, n_tiles = 0;
  int n_res = 0, box_cols = 64;
  bool gpair = false;     // conv_igemm_kernel_t<true>: CTA pairs (TMA-fed layer without residual); weight map box = bn / 2 rows
  bool strip = false;     // stride-1 k x k layer run by conv_strip_kernel (halo patch + shifted descriptors)
  int Wbox = 0, R = 0, Hbox = 0, row_blocks = 0, a_stage_bytes = 0, n_acc = 2;
  bool pair = false;      // strip conv run by CTA pairs (cta_group::2): weight map box = cout / 2 rows
  bool flat = false;      // strip conv with flat 128-position tiles (the destination buffer is pitched to Wbox)
  size_t bias16_off = 0;  // byte offset of this op's [cout] x {hi, lo, 0 x 6} fp16 bias rows in d_bias16
  size_t smem = 0;
  double flops_per_image = 0;
};

// N tile for one layer at one batch size.  Large layers take the widest tile (fewest re-reads of the activation
// K-blocks); layers with few M tiles are cut along N until the persistent grid covers the SMs.  Cost model in
// SM cycles per tile: main loop = nkb * max(MMA, L2->smem fill at ~32 B/cycle/SM), epilogue ~120 cycles per
// 16-column chunk, ~1200 cycles fixed; total = waves * tile.
static int pick_bn_tile(int cout, int m_tiles, int nkb, int sms, bool residual) {
  int best = 16;
  double best_cost = 1e30;
  for (int d = 16; d <= 256 && d <= cout; d += 16) {
    if (cout % d) continue;
    if (residual && (d % 64 || d > 128)) continue;                    // residual K-blocks are 64 columns wide; 128 beat 256 in the
                                                                      // sweep (tools/tune_bn.py): one staging pass per tile
    const long long tiles = (long long)m_tiles * (cout / d);
    const long long waves = (tiles + sms - 1) / sms;
    const double fill = (16384.0 + 128.0 * d) / 32.0;
    const int nk = nkb + (residual ? d / 64 : 0);
    const double tile = nk * std::max(2.0 * d, fill) + 1200.0 + (d / 16) * 60.0;
    const double cost = waves * tile;
    if (cost < best_cost * 0.999 || (cost <= best_cost * 1.001 && d > best)) { best_cost = cost; best = d; }
  }
  return best;
}

static FastDiv make_fastdiv(int d) {
  FastDiv f;
  int s = 0;
  while ((1ll << s) < d) ++s;
  f.sh = 31 + s;
  f.mul = static_cast<uint32_t>(((1ull << f.sh) + d - 1) / static_cast<unsigned long long>(d));
  return f;
}

// The ten Block17 blocks as one launch (block17_fused.cuh): ops [first_op, first_op + 4 * n_blocks) of the plan.
struct Fused17 {
  int first_op = -1, n_blocks = 0;
  uint8_t* d_stream = nullptr;      // n_blocks x 84 units of 16 KB: the weights in consumption order, as swizzled smem images
  float* d_bias = nullptr;          // n_blocks x 1408 fp32
  long long* d_trace = nullptr;     // FIRE_B200_TRACE17=1
  B17Params prm;
};

// The five Block35 blocks as one launch (block35_fused.cuh): ops [first_op, first_op + 4 * n_blocks).
struct Fused35 {
  int first_op = -1, n_blocks = 0;
  uint8_t* d_stream = nullptr;      // n_blocks x 19 units of 16 KB
  float* d_bias = nullptr;          // n_blocks x 448 fp32
  long long* d_trace = nullptr;     // FIRE_B200_TRACE35=1
  int* d_flags = nullptr;           // [n_blocks][flags_cap]: epoch of the last published y of (block, image) - balanced schedule only
  int flags_cap = 0, epoch = 0;
  bool balance = true;              // FIRE_B200_B35_BALANCE=0: a CTA takes whole images through all blocks
  bool coop = true;                 // cooperative launch when the chains wander: the grid only starts when all of it can be resident (FIRE_B200_B35_COOP=0: plain launch)
  B35Params prm;
};

// The 1x3 -> 3x1 -> up tail of every Block8 block as one launch (block8_fused.cuh): ops first_op + 4 j + {1, 2, 3}; the
// heads conv (first_op + 4 j) stays a conv_igemm launch.
struct Fused8 {
  int first_op = -1, n_blocks = 0;
  int gap_op = -1;                  // index of the OP_GAP that the last block's launch also computes (-1: none)
  uint8_t* d_w = nullptr;           // per block: 36 units of 12288 B (1x3, 3x1) + 7 x 12 units of 16 KB (up, by N tile)
  float* d_bias = nullptr;          // per block: [192 | 192 | 1792] fp32
  long long* d_trace = nullptr;     // FIRE_B200_TRACE8=1
  B8Params prm[B8_MAX_BLOCKS];
};
constexpr size_t B8_W_PER_BLOCK = (size_t)B8_WMID_UNITS * B8_WMID_BYTES + (size_t)B8_NTILES * B8_WUP_UNITS * B8_UNIT;

// MaxPool_3a + Conv2d_3b as one launch (pool_conv_fused.cuh): ops pool_op and pool_op + 1.
struct FusedPoolConv {
  int pool_op = -1;
  float* d_bias = nullptr;          // [80] fp32
  PoolConvParams prm;
};

struct fire_net {
  BlobHeader hdr;
  int device = 0;          // the device that was current at fire_facenet_create; every entry point runs there
  Fused17 f17;
  Fused35 f35;
  Fused8 f8;
  FusedPoolConv fpc;
  std::vector<BlobBuf> bufs;
  std::vector<OpRt> ops;
  uint8_t* d_weights = nullptr;
  uint8_t* d_bias16 = nullptr;   // per conv op: [cout] x {fp16 hi, fp16 lo, 0 x 6} of the fp32 bias (the bias MMA operand)
  // cache key of the activation tensor maps
  const void* key_in = nullptr; const void* key_ws = nullptr; const void* key_out = nullptr; int key_B = 0;
  double flops_per_image = 0;
  bool pdl = true;        // programmatic dependent launch between conv layers (FIRE_B200_PDL=0 disables)
  bool gather_l1 = false; // cp.async.ca instead of .cg for the A gather (FIRE_B200_GATHER_L1=1)
  int max_stages = 8;
  int n_issuers = CONV_MAX_ISSUERS;   // TMA issuing threads per CTA in 1x1 layers (FIRE_B200_ISSUERS=1|2|4)
  int strip_mma_warps = STRIP_MMA_WARPS;   // FIRE_B200_STRIP_MMAW=1|2|4: MMA issuing warps of conv_strip_kernel
  bool igemm_pair = true;   // FIRE_B200_IGEMM_PAIR=0: no CTA pairs in conv_igemm_kernel
  int pair_slack = 0;       // FIRE_B200_PAIR_SLACK=n: accept layers with n pairs fewer than SMs / 2
  bool strip_pair = false;  // FIRE_B200_STRIP_PAIR=1: CTA pairs (cta_group::2) in conv_strip_kernel - correct, but measured slower
  bool use_strip = true;    // FIRE_B200_STRIP=0 forces the gather path for every k x k layer (A/B experiments)
  bool trace_all = false;   // FIRE_B200_TRACE_ALL=1: forward() records every conv's timeline, synchronises and prints it
  long long* d_trace = nullptr; int trace_op = -1;   // FIRE_B200_TRACE_OP=<op index>: in-kernel timeline of that op (profile only)
  // Max-pools whose input was produced several ops earlier (the pool branches of Mixed_6a / Mixed_7a) run on a side stream,
  // concurrently with the branch convolutions that read the same tensor (FIRE_B200_SIDE_POOLS=0: in plan order)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<int> hoist_at;   // per op: index of the max-pool to launch on the side stream just before this op (-1: none)
  std::vector<char> hoisted;   // per op: this max-pool is launched early (skip it at its own position, join after it)
  int dbg_flags = 0;      // always 0 unless built with -DFIRE_B200_SKIP_EXPERIMENTS (then FIRE_B200_DBG: 1 = no gather copies, 2 = no epilogue stores, 4 = no MMA)
};

static inline int buf_wp(const BlobBuf& b) { return b.Wp > 0 ? b.Wp : b.W; }

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}


// ---- Block17 fusion: pattern match on the plan and host-side repacking of the weights ----------------------------
// unit images: byte offset of element (row n, k) inside a [128 x 64] SWIZZLE_128B / [256 x 32] SWIZZLE_64B operand
static void b17_put_sw128(uint16_t* unit, const uint16_t* W, int ldw, int row0, int k0) {
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 64; ++k)
      unit[(n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) / 2] = W[(size_t)(row0 + n) * ldw + k0 + k];
}
static void b17_put_sw64(uint16_t* unit, const uint16_t* W, int ldw, int row0, int k0) {
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < 32; ++k)
      unit[(n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2) / 2] = W[(size_t)(row0 + n) * ldw + k0 + k];
}
static bool b17_match(const std::vector<BlobOp>& ops, const std::vector<BlobBuf>& bufs, size_t i) {
  if (i + 3 >= ops.size()) return false;
  const BlobOp &h = ops[i], &a = ops[i + 1], &b = ops[i + 2], &u = ops[i + 3];
  auto conv = [](const BlobOp& o, int kh, int kw, int cin, int cout, int ph, int pw) {
    return o.kind == OP_CONV && o.kh == kh && o.kw == kw && o.stride == 1 && o.cin == cin && o.cout == cout && o.pad_h == ph && o.pad_w == pw &&
           o.H == 8 && o.W == 8 && o.Ho == 8 && o.Wo == 8 && o.k_pad == kh * kw * cin;
  };
  if (!conv(h, 1, 1, B17_C, 256, 0, 0) || !conv(a, 1, 7, 128, 128, 0, 3) || !conv(b, 7, 1, 128, 128, 3, 0) || !conv(u, 1, 1, 256, B17_C, 0, 0)) return false;
  if (h.flags != CF_RELU || a.flags != CF_RELU || b.flags != CF_RELU || u.flags != (CF_RELU | CF_RESIDUAL)) return false;
  const BlobBuf &xb = bufs[h.src_buf], &yb = bufs[u.dst_buf];
  if (xb.C != B17_C || yb.C != B17_C || h.src_coff || u.dst_coff || (xb.Wp && xb.Wp != xb.W) || (yb.Wp && yb.Wp != yb.W)) return false;
  if (u.res_buf != h.src_buf || u.res_coff) return false;
  // X = [b1a | b0 | b1c]: 1x7 reads the first 128 heads columns, `up` reads [b0 | b1c]
  if (a.src_buf != h.dst_buf || a.src_coff != h.dst_coff || b.src_buf != a.dst_buf || b.src_coff != a.dst_coff) return false;
  if (u.src_buf != h.dst_buf || u.src_coff != h.dst_coff + 128 || b.dst_buf != h.dst_buf || b.dst_coff != h.dst_coff + 256) return false;
  return true;
}
// Finds the chain and uploads its weight stream / bias table.  Returns false (fusion off) when the plan has no chain.
static bool b17_setup(fire_net* net, const std::vector<BlobOp>& ops, const uint8_t* blob, const BlobHeader& h) {
  Fused17& f = net->f17;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (!b17_match(ops, net->bufs, i)) continue;
    int n = 1;
    while (n < B17_MAX_BLOCKS && b17_match(ops, net->bufs, i + 4 * n) && ops[i + 4 * n].src_buf == ops[i + 4 * n - 1].dst_buf) ++n;
    f.first_op = (int)i; f.n_blocks = n;
    break;
  }
  if (f.first_op < 0) return false;
  std::vector<uint16_t> stream((size_t)f.n_blocks * B17_UNITS_PER_BLOCK * (B17_UNIT / 2));
  std::vector<float> bias((size_t)f.n_blocks * B17_BIAS_PER_BLOCK);
  for (int j = 0; j < f.n_blocks; ++j) {
    const BlobOp* o = &ops[f.first_op + 4 * j];
    const uint16_t* W[4];
    for (int q = 0; q < 4; ++q) W[q] = reinterpret_cast<const uint16_t*>(blob + h.weights_off + o[q].w_off);
    uint16_t* dst = stream.data() + (size_t)j * B17_UNITS_PER_BLOCK * (B17_UNIT / 2);
    auto next = [&]() { uint16_t* r = dst; dst += B17_UNIT / 2; return r; };
    for (int kb = 0; kb < 14; ++kb)
      for (int half = 0; half < 2; ++half) b17_put_sw64(next(), W[0], B17_C, 0, kb * 64 + half * 32);
    for (int q = 1; q <= 2; ++q)
      for (int s = 0; s < 7; ++s)
        for (int kb = 0; kb < 2; ++kb) b17_put_sw128(next(), W[q], 7 * 128, 0, s * 128 + kb * 64);
    for (int t = 0; t < 3; ++t)
      for (int kb = 0; kb < 4; ++kb)
        for (int half = 0; half < 2; ++half) b17_put_sw64(next(), W[3], 256, t * 256, kb * 64 + half * 32);
    for (int kb = 0; kb < 4; ++kb) b17_put_sw128(next(), W[3], 256, 768, kb * 64);
    float* bd = bias.data() + (size_t)j * B17_BIAS_PER_BLOCK;
    const int nb[4] = {256, 128, 128, B17_C};
    for (int q = 0; q < 4; ++q) {
      memcpy(bd, blob + h.weights_off + o[q].b_off, sizeof(float) * nb[q]);
      bd += nb[q];
    }
  }
  if (cudaMalloc(&f.d_stream, stream.size() * 2) != cudaSuccess || cudaMalloc(&f.d_bias, bias.size() * 4) != cudaSuccess ||
      cudaMemcpy(f.d_stream, stream.data(), stream.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(f.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaFuncSetAttribute(block17_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B17_SMEM) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(f.d_stream); cudaFree(f.d_bias);
    f.d_stream = nullptr; f.d_bias = nullptr; f.first_op = -1; f.n_blocks = 0;
    return false;
  }
  if (const char* e = getenv("FIRE_B200_TRACE17")) {
    if (e[0] == '1') {
      cudaMalloc(&f.d_trace, (size_t)148 * B17_MAX_BLOCKS * B17_TRACE_SLOTS * 8);
      cudaMemset(f.d_trace, 0, (size_t)148 * B17_MAX_BLOCKS * B17_TRACE_SLOTS * 8);
    }
  }
  return true;
}
static inline bool in_f17(const fire_net* net, size_t i) {
  return net->f17.first_op >= 0 && (int)i >= net->f17.first_op && (int)i < net->f17.first_op + 4 * net->f17.n_blocks;
}
static int run_f17(fire_net* net, int B, cudaStream_t st, bool pdl) {
  Fused17& f = net->f17;
  f.prm.pdl = pdl ? 1 : 0;
  f.prm.dbg = net->dbg_flags >> 16;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min(f.prm.n_tiles, device_sm_count()));
  cfg.blockDim = dim3(B17_THREADS);
  cfg.dynamicSmemBytes = B17_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  (void)B;
  FIRE_CUDA(cudaLaunchKernelEx(&cfg, block17_fused_kernel, f.prm));
  count_launch();
  return FIRE_OK;
}

// ---- Block35 fusion ------------------------------------------------------------------------------------------------
// generic operand images: `rows` x 64 K (128-byte rows, SWIZZLE_128B) / `rows` x 32 K (64-byte rows, SWIZZLE_64B);
// kmap (optional) = source K column of our K column
static void b35_put_sw128(uint16_t* unit, const uint16_t* W, int ldw, int row0, int rows, int k0, const int* kmap = nullptr) {
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < 64; ++k)
      unit[(n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) / 2] = W[(size_t)(row0 + n) * ldw + (kmap ? kmap[k0 + k] : k0 + k)];
}
static void b35_put_sw64(uint16_t* unit, const uint16_t* W, int ldw, int row0, int rows, int k0, const int* kmap = nullptr) {
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < 32; ++k)
      unit[(n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2) / 2] = W[(size_t)(row0 + n) * ldw + (kmap ? kmap[k0 + k] : k0 + k)];
}
static bool b35_match(const std::vector<BlobOp>& ops, const std::vector<BlobBuf>& bufs, size_t i) {
  if (i + 3 >= ops.size()) return false;
  const BlobOp &h = ops[i], &a = ops[i + 1], &b = ops[i + 2], &u = ops[i + 3];
  auto conv = [](const BlobOp& o, int k, int cin, int cout) {
    return o.kind == OP_CONV && o.kh == k && o.kw == k && o.stride == 1 && o.cin == cin && o.cout == cout && o.pad_h == k / 2 && o.pad_w == k / 2 &&
           o.H == 17 && o.W == 17 && o.Ho == 17 && o.Wo == 17 && o.k_pad == (k * k * cin + 63) / 64 * 64;
  };
  if (!conv(h, 1, B35_C, 96) || !conv(a, 3, 64, 64) || !conv(b, 3, 32, 32) || !conv(u, 1, 96, B35_C)) return false;
  if (h.flags != CF_RELU || a.flags != CF_RELU || b.flags != CF_RELU || u.flags != (CF_RELU | CF_RESIDUAL)) return false;
  const BlobBuf &xb = bufs[h.src_buf], &yb = bufs[u.dst_buf];
  if (xb.C != B35_C || yb.C != B35_C || h.src_coff || u.dst_coff || (xb.Wp && xb.Wp != xb.W) || (yb.Wp && yb.Wp != yb.W)) return false;
  if (u.res_buf != h.src_buf || u.res_coff) return false;
  // X = [b1a | b2a | b0 | b2c | b1b | b2b] (netplan._block35): conv1 X[0:64] -> X[128:192], conv2 X[160:192] -> X[96:128], up reads X[64:160]
  const int X = h.dst_buf, c0 = h.dst_coff;
  if (a.src_buf != X || a.src_coff != c0 || a.dst_buf != X || a.dst_coff != c0 + 128) return false;
  if (b.src_buf != X || b.src_coff != c0 + 160 || b.dst_buf != X || b.dst_coff != c0 + 96) return false;
  if (u.src_buf != X || u.src_coff != c0 + 64) return false;
  return true;
}
static bool b35_setup(fire_net* net, const std::vector<BlobOp>& ops, const uint8_t* blob, const BlobHeader& h) {
  Fused35& f = net->f35;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (!b35_match(ops, net->bufs, i)) continue;
    int n = 1;
    while (n < B35_MAX_BLOCKS && b35_match(ops, net->bufs, i + 4 * n) && ops[i + 4 * n].src_buf == ops[i + 4 * n - 1].dst_buf) ++n;
    f.first_op = (int)i; f.n_blocks = n;
    break;
  }
  if (f.first_op < 0) return false;
  std::vector<uint16_t> stream((size_t)f.n_blocks * B35_WSLOTS_PER_BLOCK * (B35_UNIT / 2), 0);
  std::vector<float> bias((size_t)f.n_blocks * B35_BIAS_PER_BLOCK);
  // the plan's `up` reads [b0 | b2c | b1b]; the kernel's A operand is [b0 | b1b | b2c]
  int kmap[96];
  for (int k = 0; k < 32; ++k) { kmap[k] = k; kmap[32 + k] = 64 + k; kmap[64 + k] = 32 + k; }
  for (int j = 0; j < f.n_blocks; ++j) {
    const BlobOp* o = &ops[f.first_op + 4 * j];
    const uint16_t* W[4];
    for (int q = 0; q < 4; ++q) W[q] = reinterpret_cast<const uint16_t*>(blob + h.weights_off + o[q].w_off);
    uint16_t* base = stream.data() + (size_t)j * B35_WSLOTS_PER_BLOCK * (B35_UNIT / 2);
    auto unit = [&](int slot) { return base + (size_t)slot * (B35_UNIT / 2); };
    for (int kb = 0; kb < 4; ++kb) b35_put_sw128(unit(kb), W[0], o[0].k_pad, 0, 96, kb * 64);
    for (int t = 0; t < 9; ++t) b35_put_sw128(unit(4 + t), W[1], o[1].k_pad, 0, 64, t * 64);
    for (int t = 0; t < 9; ++t) b35_put_sw64(unit(t < 5 ? 13 : 14) + (size_t)(t < 5 ? t : t - 5) * 1024, W[2], o[2].k_pad, 0, 32, t * 32);
    for (int nh = 0; nh < 2; ++nh) {
      b35_put_sw128(unit(15 + 2 * nh), W[3], o[3].k_pad, nh * 128, 128, 0, kmap);
      b35_put_sw64(unit(16 + 2 * nh), W[3], o[3].k_pad, nh * 128, 128, 64, kmap);
    }
    float* bd = bias.data() + (size_t)j * B35_BIAS_PER_BLOCK;
    const int nb[4] = {96, 64, 32, B35_C};
    for (int q = 0; q < 4; ++q) {
      memcpy(bd, blob + h.weights_off + o[q].b_off, sizeof(float) * nb[q]);
      bd += nb[q];
    }
  }
  if (cudaMalloc(&f.d_stream, stream.size() * 2) != cudaSuccess || cudaMalloc(&f.d_bias, bias.size() * 4) != cudaSuccess ||
      cudaMemcpy(f.d_stream, stream.data(), stream.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(f.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaFuncSetAttribute(block35_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B35_SMEM) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(f.d_stream); cudaFree(f.d_bias);
    f.d_stream = nullptr; f.d_bias = nullptr; f.first_op = -1; f.n_blocks = 0;
    return false;
  }
  if (const char* e = getenv("FIRE_B200_B35_BALANCE")) f.balance = e[0] != '0';
  if (const char* e = getenv("FIRE_B200_B35_COOP")) f.coop = e[0] != '0';
  if (const char* e = getenv("FIRE_B200_TRACE35")) {
    if (e[0] == '1') {
      cudaMalloc(&f.d_trace, (size_t)148 * 16 * 24 * 8);
      cudaMemset(f.d_trace, 0, (size_t)148 * 16 * 24 * 8);
    }
  }
  return true;
}
static inline bool in_f35(const fire_net* net, size_t i) {
  return net->f35.first_op >= 0 && (int)i >= net->f35.first_op && (int)i < net->f35.first_op + 4 * net->f35.n_blocks;
}
static int run_f35(fire_net* net, cudaStream_t st, bool pdl) {
  Fused35& f = net->f35;
  f.prm.pdl = pdl ? 1 : 0;
  f.epoch = (int)((unsigned)f.epoch + 1u);            // launch number (wraps; the kernel compares differences): the flags are never reset
  f.prm.epoch = f.epoch;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min(f.prm.n_images, device_sm_count()));
  cfg.blockDim = dim3(B35_THREADS);
  cfg.dynamicSmemBytes = B35_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  // Wandering chains wait on flags written by other CTAs of this grid: ask for a COOPERATIVE launch, which only starts the grid
  // when all of its CTAs can be resident at once (measured cost: 0.8 us per forward; FIRE_B200_B35_COOP=0 turns it off)
  const bool wander = f.balance && (f.prm.n_images % (int)cfg.gridDim.x) != 0;
  if (wander && f.coop) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, block35_fused_kernel, f.prm);
  if (e != cudaSuccess && wander && f.coop) {            // not launchable that way here: plain launch from now on
    cudaGetLastError();
    f.coop = false;
    cfg.numAttrs = pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, block35_fused_kernel, f.prm);
  }
  if (e != cudaSuccess) return fail(FIRE_ERR_CUDA, "launch of block35_fused_kernel failed: %s", cudaGetErrorString(e));
  count_launch();
  return FIRE_OK;
}

// ---- Block8 tail fusion ---------------------------------------------------------------------------------------------
static bool b8_match(const std::vector<BlobOp>& ops, const std::vector<BlobBuf>& bufs, size_t i) {
  if (i + 3 >= ops.size()) return false;
  const BlobOp &h = ops[i], &a = ops[i + 1], &b = ops[i + 2], &u = ops[i + 3];
  auto conv = [](const BlobOp& o, int kh, int kw, int cin, int cout, int ph, int pw) {
    return o.kind == OP_CONV && o.kh == kh && o.kw == kw && o.stride == 1 && o.cin == cin && o.cout == cout && o.pad_h == ph && o.pad_w == pw &&
           o.H == 3 && o.W == 3 && o.Ho == 3 && o.Wo == 3 && o.k_pad == kh * kw * cin;
  };
  if (!conv(h, 1, 1, B8_C, 2 * B8_MID, 0, 0) || !conv(a, 1, 3, B8_MID, B8_MID, 0, 1) || !conv(b, 3, 1, B8_MID, B8_MID, 1, 0) ||
      !conv(u, 1, 1, 2 * B8_MID, B8_C, 0, 0)) return false;
  if (h.flags != CF_RELU || a.flags != CF_RELU || b.flags != CF_RELU || (u.flags != (CF_RELU | CF_RESIDUAL) && u.flags != CF_RESIDUAL)) return false;
  const BlobBuf &xb = bufs[h.src_buf], &yb = bufs[u.dst_buf], &Xb = bufs[h.dst_buf];
  if (xb.C != B8_C || yb.C != B8_C || h.src_coff || u.dst_coff || (xb.Wp && xb.Wp != xb.W) || (yb.Wp && yb.Wp != yb.W) || (Xb.Wp && Xb.Wp != Xb.W)) return false;
  if (u.res_buf != h.src_buf || u.res_coff) return false;
  // X = [b1a | b0 | b1c]: 1x3 reads the first 192 heads columns, `up` reads [b0 | b1c]
  if (a.src_buf != h.dst_buf || a.src_coff != h.dst_coff || b.src_buf != a.dst_buf || b.src_coff != a.dst_coff) return false;
  if (u.src_buf != h.dst_buf || u.src_coff != h.dst_coff + B8_MID || b.dst_buf != h.dst_buf || b.dst_coff != h.dst_coff + 2 * B8_MID) return false;
  return true;
}
static bool b8_setup(fire_net* net, const std::vector<BlobOp>& ops, const uint8_t* blob, const BlobHeader& h) {
  Fused8& f = net->f8;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (!b8_match(ops, net->bufs, i)) continue;
    int n = 1;
    while (n < B8_MAX_BLOCKS && b8_match(ops, net->bufs, i + 4 * n)) ++n;
    f.first_op = (int)i; f.n_blocks = n;
    break;
  }
  if (f.first_op < 0) return false;
  {
    // global average pool right after the last block, over the whole of its output: folded into that block's launch
    const size_t g = (size_t)f.first_op + 4 * f.n_blocks;
    const BlobOp& u = ops[g - 1];
    if (g < ops.size() && ops[g].kind == OP_GAP && ops[g].src_buf == u.dst_buf && ops[g].src_coff == 0 && ops[g].cin == B8_C && ops[g].H == 3 &&
        ops[g].W == 3 && net->bufs[ops[g].dst_buf].elt == 2 && !(getenv("FIRE_B200_FUSE_GAP") && getenv("FIRE_B200_FUSE_GAP")[0] == '0'))
      f.gap_op = (int)g;
  }
  std::vector<uint16_t> w((size_t)f.n_blocks * B8_W_PER_BLOCK / 2, 0);
  std::vector<float> bias((size_t)f.n_blocks * B8_BIAS_PER_BLOCK);
  for (int j = 0; j < f.n_blocks; ++j) {
    const BlobOp* o = &ops[f.first_op + 4 * j];
    const uint16_t* W[4];
    for (int q = 0; q < 4; ++q) W[q] = reinterpret_cast<const uint16_t*>(blob + h.weights_off + o[q].w_off);
    uint16_t* dst = w.data() + (size_t)j * B8_W_PER_BLOCK / 2;
    // 1x3 then 3x1: (tap, 64-channel K-block, 32-channel half) -> [192 x 32] SWIZZLE_64B images
    for (int q = 1; q <= 2; ++q)
      for (int t = 0; t < 3; ++t)
        for (int kb = 0; kb < 3; ++kb)
          for (int half = 0; half < 2; ++half) { b35_put_sw64(dst, W[q], 3 * B8_MID, 0, B8_MID, t * B8_MID + kb * 64 + half * 32); dst += B8_WMID_BYTES / 2; }
    // up: per N tile, (K-block, half) -> [256 x 32] SWIZZLE_64B images
    for (int nt = 0; nt < B8_NTILES; ++nt)
      for (int kb = 0; kb < 6; ++kb)
        for (int half = 0; half < 2; ++half) { b35_put_sw64(dst, W[3], 2 * B8_MID, nt * B8_NT, B8_NT, kb * 64 + half * 32); dst += B8_UNIT / 2; }
    float* bd = bias.data() + (size_t)j * B8_BIAS_PER_BLOCK;
    memcpy(bd, blob + h.weights_off + o[1].b_off, sizeof(float) * B8_MID);
    memcpy(bd + B8_MID, blob + h.weights_off + o[2].b_off, sizeof(float) * B8_MID);
    memcpy(bd + 2 * B8_MID, blob + h.weights_off + o[3].b_off, sizeof(float) * B8_C);
  }
  if (cudaMalloc(&f.d_w, w.size() * 2) != cudaSuccess || cudaMalloc(&f.d_bias, bias.size() * 4) != cudaSuccess ||
      cudaMemcpy(f.d_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(f.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaFuncSetAttribute(block8_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B8_SMEM) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(f.d_w); cudaFree(f.d_bias);
    f.d_w = nullptr; f.d_bias = nullptr; f.first_op = -1; f.n_blocks = 0; f.gap_op = -1;
    return false;
  }
  if (const char* e = getenv("FIRE_B200_TRACE8")) {
    if (e[0] == '1') {
      cudaMalloc(&f.d_trace, (size_t)B8_MAX_BLOCKS * 148 * B8_TRACE_SLOTS * 8);
      cudaMemset(f.d_trace, 0, (size_t)B8_MAX_BLOCKS * 148 * B8_TRACE_SLOTS * 8);
    }
  }
  return true;
}
// role of op i: -1 = not part of a fused Block8 tail, 0 = the block's heads conv (runs as itself), 1 = the op that launches
// the fused tail (the 1x3 conv's slot), 2 = covered by that launch
static inline int f8_role(const fire_net* net, size_t i) {
  const Fused8& f = net->f8;
  if (f.first_op < 0 || (int)i < f.first_op || (int)i >= f.first_op + 4 * f.n_blocks) return -1;
  const int k = ((int)i - f.first_op) & 3;
  return k == 0 ? 0 : k == 1 ? 1 : 2;
}
static int run_f8(fire_net* net, size_t i, cudaStream_t st, bool pdl) {
  Fused8& f = net->f8;
  B8Params& prm = f.prm[((int)i - f.first_op) >> 2];
  prm.pdl = pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(prm.n_groups * B8_NTILES));
  cfg.blockDim = dim3(B8_THREADS);
  cfg.dynamicSmemBytes = B8_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  FIRE_CUDA(cudaLaunchKernelEx(&cfg, block8_fused_kernel, prm));
  count_launch();
  return FIRE_OK;
}

// ---- stem max-pool + 1x1 conv fusion ---------------------------------------------------------------------------------
static bool pc_setup(fire_net* net, const std::vector<BlobOp>& ops, const uint8_t* blob, const BlobHeader& h) {
  FusedPoolConv& f = net->fpc;
  for (size_t i = 0; i + 1 < ops.size(); ++i) {
    const BlobOp &po = ops[i], &co = ops[i + 1];
    if (po.kind != OP_MAXPOOL || po.cin != PC_CIN || po.H != PC_IN || po.W != PC_IN || po.Ho != PC_OUT || po.Wo != PC_OUT || po.src_coff || po.dst_coff) continue;
    if (co.kind != OP_CONV || co.kh != 1 || co.kw != 1 || co.stride != 1 || co.cin != PC_CIN || co.cout != PC_COUT || co.k_pad != PC_CIN ||
        co.flags != CF_RELU || co.src_buf != po.dst_buf || co.src_coff || co.dst_coff) continue;
    const BlobBuf &sb = net->bufs[po.src_buf], &pb = net->bufs[po.dst_buf], &db = net->bufs[co.dst_buf];
    if (sb.C != PC_CIN || pb.C != PC_CIN || db.C != PC_COUT || (db.Wp && db.Wp != db.W) || sb.external || db.external) continue;
    bool other_reader = false;                          // the pooled tensor must have no other consumer: it is never written
    for (size_t j = 0; j < ops.size(); ++j)
      if (j != i + 1 && (ops[j].src_buf == po.dst_buf || ops[j].res_buf == po.dst_buf)) other_reader = true;
    if (other_reader) continue;
    f.pool_op = (int)i;
    break;
  }
  if (f.pool_op < 0) return false;
  const BlobOp& co = ops[f.pool_op + 1];
  if (cudaMalloc(&f.d_bias, sizeof(float) * PC_COUT) != cudaSuccess ||
      cudaMemcpy(f.d_bias, blob + h.weights_off + co.b_off, sizeof(float) * PC_COUT, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaFuncSetAttribute(pool_conv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(f.d_bias);
    f.d_bias = nullptr; f.pool_op = -1;
    return false;
  }
  return true;
}
static int run_pc(fire_net* net, cudaStream_t st, bool pdl) {
  FusedPoolConv& f = net->fpc;
  f.prm.pdl = pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min(f.prm.n_images * PC_ROW_BLOCKS, device_sm_count()));
  cfg.blockDim = dim3(PC_THREADS);
  cfg.dynamicSmemBytes = PC_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  FIRE_CUDA(cudaLaunchKernelEx(&cfg, pool_conv_fused_kernel, f.prm));
  count_launch();
  return FIRE_OK;
}

extern "C" {

int fire_facenet_create(const void* host_blob, size_t bytes, fire_net_t** out) {
  if (!host_blob || !out) return fail(FIRE_ERR_ARG, "fire_facenet_create: NULL argument");
  if (bytes < sizeof(BlobHeader)) return fail(FIRE_ERR_ARG, "fire_facenet_create: blob too small");
  const uint8_t* p = static_cast<const uint8_t*>(host_blob);
  BlobHeader h;
  memcpy(&h, p, sizeof(h));
  if (memcmp(h.magic, "FIREB200", 8) != 0 || h.version != (int32_t)BLOB_VERSION)
    return fail(FIRE_ERR_ARG, "fire_facenet_create: bad magic/version (want FIREB200 v%u)", BLOB_VERSION);
  const size_t meta = sizeof(BlobHeader) + sizeof(BlobBuf) * h.n_bufs + sizeof(BlobOp) * h.n_ops;
  if (h.n_ops <= 0 || h.n_bufs <= 0 || meta > bytes || (size_t)(h.weights_off + h.weights_bytes) > bytes)
    return fail(FIRE_ERR_ARG, "fire_facenet_create: truncated blob");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(FIRE_ERR_CUDA, "no CUDA device (fire_b200 has no CPU fallback)");
  fire_net* net = new (std::nothrow) fire_net();
  if (!net) return fail(FIRE_ERR_STATE, "out of host memory");
  net->hdr = h;
  net->device = dev;
  net->bufs.resize(h.n_bufs);
  memcpy(net->bufs.data(), p + sizeof(BlobHeader), sizeof(BlobBuf) * h.n_bufs);
  std::vector<BlobOp> ops(h.n_ops);
  memcpy(ops.data(), p + sizeof(BlobHeader) + sizeof(BlobBuf) * h.n_bufs, sizeof(BlobOp) * h.n_ops);
  cudaError_t e = cudaMalloc(&net->d_weights, h.weights_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(net->d_weights, p + h.weights_off, h.weights_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(net->d_weights);
    delete net;
    return fail(FIRE_ERR_CUDA, "fire_facenet_create: weight upload failed: %s", cudaGetErrorString(e));
  }
  {
    // bias operand of the bias MMA: per output channel {fp16 hi, fp16 lo, 0 x 6}, hi + lo ~ the fp32 bias to 2^-22
    size_t total = 0;
    for (const BlobOp& o : ops) if (o.kind == OP_CONV) total += (size_t)o.cout * 16;
    std::vector<uint16_t> tab(total / 2, 0);
    size_t off = 0;
    for (const BlobOp& o : ops) {
      if (o.kind != OP_CONV) continue;
      if ((size_t)(h.weights_off + o.b_off) + (size_t)o.cout * 4 > bytes) { cudaFree(net->d_weights); delete net; return fail(FIRE_ERR_ARG, "fire_facenet_create: bias out of range"); }
      const float* b = reinterpret_cast<const float*>(p + h.weights_off + o.b_off);
      for (int c = 0; c < o.cout; ++c) {
        float v = std::min(std::max(b[c], -60000.f), 60000.f);
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        memcpy(&tab[off / 2 + (size_t)c * 8], &hi, 2);
        memcpy(&tab[off / 2 + (size_t)c * 8 + 1], &lo, 2);
      }
      off += (size_t)o.cout * 16;
    }
    e = cudaMalloc(&net->d_bias16, std::max<size_t>(total, 16));
    if (e == cudaSuccess) e = cudaMemcpy(net->d_bias16, tab.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(net->d_weights); cudaFree(net->d_bias16);
      delete net;
      return fail(FIRE_ERR_CUDA, "fire_facenet_create: bias upload failed: %s", cudaGetErrorString(e));
    }
  }
  static bool attr_done[FIRE_MAX_DEVICES] = {};        // the shared-memory opt-in is per device (context), not per process
  if (dev < 0 || dev >= FIRE_MAX_DEVICES) { cudaFree(net->d_weights); cudaFree(net->d_bias16); delete net; return fail(FIRE_ERR_ARG, "device %d out of range", dev); }
  if (!attr_done[dev]) {
    e = cudaFuncSetAttribute(conv_igemm_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_igemm_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_strip_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_strip_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      cudaFree(net->d_weights);
      delete net;
      return fail(FIRE_ERR_CUDA, "cudaFuncSetAttribute(conv_igemm_kernel): %s", cudaGetErrorString(e));
    }
    attr_done[dev] = true;
  }
  size_t bias16_off = 0;
  for (const BlobOp& o : ops) {
    OpRt r;
    r.op = o;
    if (o.src_buf < 0 || o.src_buf >= h.n_bufs || o.dst_buf < 0 || o.dst_buf >= h.n_bufs || o.res_buf >= h.n_bufs) {
      cudaFree(net->d_weights); delete net;
      return fail(FIRE_ERR_ARG, "fire_facenet_create: op references a buffer out of range");
    }
    if (o.kind == OP_CONV) {
      if (o.cout % 16 || o.cout > CONV_MAX_COUT || o.k_pad % 64 || o.cin % 8) {
        cudaFree(net->d_weights); delete net;
        return fail(FIRE_ERR_ARG, "fire_facenet_create: conv op with unsupported shape (cout=%d k_pad=%d cin=%d)",
                    o.cout, o.k_pad, o.cin);
      }
      r.tma_a = (o.kh == 1 && o.kw == 1 && o.stride == 1 && o.pad_h == 0 && o.pad_w == 0);
      r.im2col = !r.tma_a && o.kh * o.kw > 1 && o.cin % 64 == 0 && o.k_pad == o.kh * o.kw * o.cin &&
                 !(o.flags & (CF_RESIDUAL | CF_OUT_F32));
      if ((o.flags & CF_RESIDUAL) && (!r.tma_a || o.cout % 64)) {
        cudaFree(net->d_weights); cudaFree(net->d_bias16); delete net;
        return fail(FIRE_ERR_ARG, "fire_facenet_create: residual is supported on 1x1/stride-1 convs with cout %% 64 == 0");
      }
      r.bias16_off = bias16_off;
      bias16_off += (size_t)o.cout * 16;
      r.flops_per_image = 2.0 * o.Ho * o.Wo * (double)o.cout * (o.flop_k > 0 ? o.flop_k : o.kh * o.kw * o.cin);
      net->flops_per_image += r.flops_per_image;
    }
    net->ops.push_back(r);
  }
  const char* pdl_env = getenv("FIRE_B200_PDL");
  net->pdl = !(pdl_env && pdl_env[0] == '0');
  const char* l1_env = getenv("FIRE_B200_GATHER_L1");
  net->gather_l1 = l1_env && l1_env[0] == '1';
#ifdef FIRE_B200_SKIP_EXPERIMENTS
  const char* dbg_env = getenv("FIRE_B200_DBG");
  if (dbg_env) net->dbg_flags = (atoi(dbg_env) & 7) << 16;
#endif
  const char* tr_env = getenv("FIRE_B200_TRACE_OP");
  if (tr_env && atoi(tr_env) >= 0 && atoi(tr_env) < (int)net->ops.size()) {
    net->trace_op = atoi(tr_env);
    cudaMalloc(&net->d_trace, 8 * 8 * 512);
    cudaMemset(net->d_trace, 0, 8 * 8 * 512);
  }
  if (const char* e = getenv("FIRE_B200_IM2COL")) {
    if (e[0] == '0') for (OpRt& r : net->ops) r.im2col = false;      // A/B experiments and the parity test: cp.async gather instead
  }
  if (const char* e = getenv("FIRE_B200_STRIP_MMAW")) net->strip_mma_warps = atoi(e) >= 2 ? 2 : 1;
  if (const char* e = getenv("FIRE_B200_STRIP_PAIR")) net->strip_pair = e[0] == '1';
  if (const char* e = getenv("FIRE_B200_IGEMM_PAIR")) net->igemm_pair = e[0] != '0';
  if (const char* e = getenv("FIRE_B200_PAIR_SLACK")) net->pair_slack = atoi(e);
  const char* sp_env = getenv("FIRE_B200_STRIP");
  net->use_strip = !(sp_env && sp_env[0] == '0');
  const char* ta_env = getenv("FIRE_B200_TRACE_ALL");
  if (ta_env && ta_env[0] == '1') {
    net->trace_all = true; net->trace_op = 0;
    cudaFree(net->d_trace);
    cudaMalloc(&net->d_trace, net->ops.size() * 4096 * 8);
    cudaMemset(net->d_trace, 0, net->ops.size() * 4096 * 8);
  }
  const char* is_env = getenv("FIRE_B200_ISSUERS");
  if (is_env) net->n_issuers = std::max(1, std::min(CONV_MAX_ISSUERS, atoi(is_env)));   // 1, 2 or 4 are used
  const char* st_env = getenv("FIRE_B200_MAX_STAGES");
  if (st_env) net->max_stages = std::max(3, std::min(12, atoi(st_env)));
  {
    const char* f17_env = getenv("FIRE_B200_FUSE17");
    if (!(f17_env && f17_env[0] == '0')) b17_setup(net, ops, p, h);
    const char* f35_env = getenv("FIRE_B200_FUSE35");
    if (!(f35_env && f35_env[0] == '0')) b35_setup(net, ops, p, h);
    const char* f8_env = getenv("FIRE_B200_FUSE8");
    if (!(f8_env && f8_env[0] == '0')) b8_setup(net, ops, p, h);
    const char* fp_env = getenv("FIRE_B200_FUSE_POOL");
    if (!(fp_env && fp_env[0] == '0')) pc_setup(net, ops, p, h);
  }
  {
    net->hoist_at.assign(ops.size(), -1);
    net->hoisted.assign(ops.size(), 0);
    const char* sp = getenv("FIRE_B200_SIDE_POOLS");
    if (!(sp && sp[0] == '0') && cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) == cudaSuccess) {
      for (size_t i = 0; i < ops.size(); ++i) {
        if (ops[i].kind != OP_MAXPOOL) continue;
        int prod = -1;                                   // last op before i that writes the pool's input
        for (int j = (int)i - 1; j >= 0; --j) if (ops[j].dst_buf == ops[i].src_buf) { prod = j; break; }
        if (prod < 0 || prod + 1 >= (int)i) continue;    // input of the net / produced by the previous op: nothing to overlap with
        bool ok = i + 1 < ops.size();
        for (int j = prod + 1; j < (int)i && ok; ++j) {  // the ops it overtakes must not touch its output slice
          const bool same_buf_w = ops[j].dst_buf == ops[i].dst_buf, same_buf_r = ops[j].src_buf == ops[i].dst_buf || ops[j].res_buf == ops[i].dst_buf;
          const int lo = ops[i].dst_coff, hi = lo + ops[i].cout;
          if (same_buf_w && ops[j].dst_coff < hi && ops[j].dst_coff + ops[j].cout > lo) ok = false;
          if (same_buf_r) ok = false;
        }
        if (!ok || net->hoist_at[prod + 1] >= 0) continue;
        net->hoist_at[prod + 1] = (int)i;
        net->hoisted[i] = 1;
      }
    } else {
      cudaGetLastError();
    }
  }
  *out = net;
  return FIRE_OK;
}

int fire_facenet_destroy(fire_net_t* net) {
  if (!net) return FIRE_OK;
  use_device(net->device);
  cudaFree(net->d_weights);
  cudaFree(net->d_bias16);
  cudaFree(net->d_trace);
  cudaFree(net->f17.d_stream); cudaFree(net->f17.d_bias); cudaFree(net->f17.d_trace);
  cudaFree(net->f35.d_stream); cudaFree(net->f35.d_bias); cudaFree(net->f35.d_trace); cudaFree(net->f35.d_flags);
  cudaFree(net->f8.d_w); cudaFree(net->f8.d_bias); cudaFree(net->f8.d_trace);
  cudaFree(net->fpc.d_bias);
  if (net->ev_fork) cudaEventDestroy(net->ev_fork);
  if (net->ev_join) cudaEventDestroy(net->ev_join);
  if (net->side) cudaStreamDestroy(net->side);
  delete net;
  return FIRE_OK;
}

int fire_facenet_dim(const fire_net_t* net) { return net ? net->hdr.D : 0; }
int fire_facenet_num_ops(const fire_net_t* net) { return net ? (int)net->ops.size() : 0; }
int fire_facenet_num_launches(const fire_net_t* net) {
  if (!net) return 0;
  int n = (int)net->ops.size();
  if (net->f17.first_op >= 0) n -= 4 * net->f17.n_blocks - 1;
  if (net->f35.first_op >= 0) n -= 4 * net->f35.n_blocks - 1;
  if (net->f8.first_op >= 0) n -= 2 * net->f8.n_blocks + (net->f8.gap_op >= 0 ? 1 : 0);
  if (net->fpc.pool_op >= 0) n -= 1;
  return n;
}
double fire_facenet_flops(const fire_net_t* net) { return net ? net->flops_per_image : 0.0; }

size_t fire_facenet_workspace(const fire_net_t* net, int B) {
  if (!net || B <= 0) return 0;
  return (size_t)net->hdr.ws_bytes_per_image * (size_t)B + 1024;
}

}  // extern "C"

static void* buf_ptr(const fire_net* net, int buf, int B, const void* in, void* ws, void* out_raw) {
  if (buf == net->hdr.in_buf || net->bufs[buf].external == 2) return const_cast<void*>(in);      // 2: a pixel-pair view of the network input
  if (buf == net->hdr.out_buf) return out_raw;
  return static_cast<uint8_t*>(ws) + (size_t)net->bufs[buf].offset * (size_t)B;
}

static int run_op(fire_net* net, OpRt& r, int B, const void* in, void* ws, float* out_raw, cudaStream_t st, bool pdl) {
  const BlobOp& o = r.op;
  const BlobBuf& sb = net->bufs[o.src_buf];
  const BlobBuf& db = net->bufs[o.dst_buf];
  const __half* src = static_cast<const __half*>(buf_ptr(net, o.src_buf, B, in, ws, out_raw));
  void* dst = buf_ptr(net, o.dst_buf, B, in, ws, out_raw);
  if (o.kind == OP_CONV) {
    if (r.strip) {
      StripParams q;
      q.bias16 = reinterpret_cast<const uint4*>(net->d_bias16 + r.bias16_off);
      q.cin = o.cin; q.cout = o.cout; q.kh = o.kh; q.kw = o.kw; q.pad_h = o.pad_h; q.pad_w = o.pad_w;
      q.k16_steps = o.kh * o.kw * o.cin / 16; q.nkb = o.k_pad / 64;
      q.Wbox = r.Wbox; q.R = r.R; q.Hbox = r.Hbox; q.row_blocks = r.row_blocks;
      q.pair = r.pair ? 1 : 0;
      q.total_tiles = (r.pair ? (B + 1) / 2 : B) * r.row_blocks;          // pair mode: one scheduling unit = the same position block of two images
      q.flat = r.flat ? 1 : 0; q.Ho = o.Ho; q.d_wbox = make_fastdiv(r.Wbox);
      q.a_stage_bytes = r.a_stage_bytes; q.stages = r.stages; q.tmem_cols = r.tmem_cols; q.flags = o.flags | net->dbg_flags;
      if (net->d_trace && !net->trace_all) q.flags |= CF_DBG_PHASES;
      q.pdl = pdl ? 1 : 0; q.box_cols = r.box_cols; q.n_acc = r.n_acc;
      // two issuing warps pay off when the main loop is long enough to be issue-bound (Conv2d_2a / 2b: 18 K steps; 1a has 4)
      q.n_mma_warps = (q.k16_steps >= 8 && r.n_acc >= 4) ? std::min(net->strip_mma_warps, 2) : 1;
      q.trace = !net->d_trace ? nullptr : net->trace_all ? net->d_trace + (size_t)(&r - net->ops.data()) * 4096
                : (&r == &net->ops[net->trace_op] ? net->d_trace : nullptr);
      q.d_rowblocks = make_fastdiv(r.row_blocks);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = r.pair ? dim3((unsigned)std::min<long long>(2LL * q.total_tiles, device_sm_count() & ~1))
                           : dim3((unsigned)std::min<long long>((long long)B * r.row_blocks, device_sm_count()));
      cfg.blockDim = dim3(STRIP_THREADS);
      cfg.dynamicSmemBytes = r.smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      int na = 0;
      if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      if (r.pair) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
      }
      cfg.attrs = attr;
      cfg.numAttrs = na;
      if (r.pair) FIRE_CUDA(cudaLaunchKernelEx(&cfg, conv_strip_kernel_t<true>, r.tmap_w, r.tmap_a, r.tmap_out, q));
      else FIRE_CUDA(cudaLaunchKernelEx(&cfg, conv_strip_kernel_t<false>, r.tmap_w, r.tmap_a, r.tmap_out, q));
      count_launch();
      return FIRE_OK;
    }
    ConvParams p;
    p.in = src; p.in_ld = sb.C; p.in_coff = o.src_coff;
    p.out = dst; p.out_ld = db.C; p.out_coff = o.dst_coff;
    p.bias16 = reinterpret_cast<const uint4*>(net->d_bias16 + r.bias16_off);
    p.H = o.H; p.W = o.W; p.Ho = o.Ho; p.Wo = o.Wo; p.kh = o.kh; p.kw = o.kw; p.stride = o.stride;
    p.pad_h = o.pad_h; p.pad_w = o.pad_w; p.cin = o.cin; p.cout = o.cout; p.k_real = o.kh * o.kw * o.cin;
    p.nkb = o.k_pad / 64; p.flags = o.flags; p.bn_tile = r.bn_tile; p.M_total = B * o.Ho * o.Wo;
    p.stages = r.stages; p.tma_a = r.tma_a ? 1 : (r.im2col ? 2 : 0); p.tmem_cols = r.tmem_cols;
    p.cpb = std::max(1, o.cin / 64); p.d_cpb = make_fastdiv(p.cpb);
    p.m_tiles = r.m_tiles; p.n_tiles = r.n_tiles; p.pdl = pdl ? 1 : 0;
    p.flags |= net->dbg_flags;
    if (net->d_trace && !net->trace_all) p.flags |= CF_DBG_PHASES;
    p.n_res = r.n_res; p.box_cols = r.box_cols;
    p.n_issuers = r.n_issuers;
    p.trace = !net->d_trace ? nullptr : net->trace_all ? net->d_trace + (size_t)(&r - net->ops.data()) * 4096
              : (&r == &net->ops[net->trace_op] ? net->d_trace : nullptr);
    p.d_howo = make_fastdiv(o.Ho * o.Wo); p.d_wo = make_fastdiv(o.Wo); p.d_cin = make_fastdiv(o.cin); p.d_kw = make_fastdiv(o.kw);
    p.d_ntiles = make_fastdiv(r.n_tiles);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = r.gpair ? dim3((unsigned)std::min<long long>(2LL * ((r.m_tiles + 1) / 2) * r.n_tiles, device_sm_count() & ~1))
                          : dim3((unsigned)std::min<long long>((long long)r.m_tiles * r.n_tiles, device_sm_count()));
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = r.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (r.gpair) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (r.gpair)
      FIRE_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel_t<true>, r.tmap_w, (r.tma_a || r.im2col) ? r.tmap_a : r.tmap_w, r.tmap_w, r.tmap_out, p));
    else
      FIRE_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel_t<false>, r.tmap_w, (r.tma_a || r.im2col) ? r.tmap_a : r.tmap_w, r.n_res ? r.tmap_res : r.tmap_w,
                                   (o.flags & CF_OUT_F32) ? r.tmap_w : r.tmap_out, p));
  } else if (o.kind == OP_MAXPOOL) {
    const long long total = (long long)B * o.Ho * o.Wo * (o.cin / 8);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    maxpool3x3s2_kernel<<<blocks, 256, 0, st>>>(src, sb.C, o.src_coff, static_cast<__half*>(dst), db.C, o.dst_coff, B, o.H,
                                                buf_wp(sb), o.Ho, o.Wo, o.cin);      // reads a pitched input through its pitch
    FIRE_LAUNCH_CHECK("maxpool3x3s2_kernel");
  } else if (o.kind == OP_GAP) {
    const int n = B * (o.cin / 2);
    gap_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, sb.C, o.src_coff, static_cast<__half*>(dst), db.C, B, o.H * o.W, o.cin);
    FIRE_LAUNCH_CHECK("gap_kernel");
  } else {
    return fail(FIRE_ERR_ARG, "unknown op kind %d", o.kind);
  }
  count_launch();
  return FIRE_OK;
}

static int prepare(fire_net* net, const void* in, int B, float* out_raw, void* ws, size_t ws_bytes) {
  if (!net || !in || !out_raw || !ws) return fail(FIRE_ERR_ARG, "fire_facenet_forward: NULL argument");
  if (B <= 0 || B > 4096) return fail(FIRE_ERR_UNSUPPORTED, "fire_facenet_forward: B=%d outside [1,4096]", B);
  if (ws_bytes < fire_facenet_workspace(net, B))
    return fail(FIRE_ERR_ARG, "fire_facenet_forward: workspace %zu < required %zu", ws_bytes, fire_facenet_workspace(net, B));
  if ((reinterpret_cast<uintptr_t>(ws) & 255) || (reinterpret_cast<uintptr_t>(in) & 15))
    return fail(FIRE_ERR_ARG, "fire_facenet_forward: workspace must be 256-byte and input 16-byte aligned");
  if (net->key_in != in || net->key_ws != ws || net->key_out != out_raw || net->key_B != B) {
    const int sms = device_sm_count();
    for (OpRt& r : net->ops) {
      if (r.op.kind != OP_CONV) continue;
      const BlobOp& o = r.op;
      r.m_tiles = (B * o.Ho * o.Wo + CONV_BM - 1) / CONV_BM;
      r.strip = false;
      {
        // strip mode: stride-1 k x k, one swizzle-wide channel panel, weights resident in shared memory
        const int wbox = o.Wo + o.kw - 1;
        const bool shape_ok = net->use_strip && o.stride == 1 && o.kh * o.kw > 1 && (o.cin == 16 || o.cin == 32 || o.cin == 64) &&
                              o.cout <= 256 && !(o.flags & (CF_RESIDUAL | CF_OUT_F32)) && wbox <= CONV_BM &&
                              (size_t)(o.k_pad / 64) * o.cout * 128 <= 96 * 1024 && (o.kh * o.kw * o.cin) % 16 == 0;
        const BlobBuf& sbuf = net->bufs[o.src_buf];
        const BlobBuf& dbuf = net->bufs[o.dst_buf];
        if (!shape_ok && (buf_wp(dbuf) != dbuf.W || buf_wp(sbuf) != sbuf.W))
          return fail(FIRE_ERR_UNSUPPORTED, "op %d touches a pitched buffer but cannot run in strip mode (FIRE_B200_STRIP=0 needs Plan(pitched=False))",
                      (int)(&r - net->ops.data()));
        if (shape_ok) {
          r.strip = true;
          r.Wbox = wbox;
          r.flat = buf_wp(dbuf) != dbuf.W;
          if (r.flat && buf_wp(dbuf) != wbox)
            return fail(FIRE_ERR_ARG, "pitched destination of op %d must have pitch Wo + kw - 1 = %d", (int)(&r - net->ops.data()), wbox);
          int rows_alloc;
          if (r.flat) {
            const int rows_out = (wbox - 1 + CONV_BM - 1) / wbox + 1;          // output rows a 128-position tile can touch
            r.R = 0;
            r.Hbox = rows_out + o.kh - 1;
            r.row_blocks = (o.Ho * wbox + CONV_BM - 1) / CONV_BM;
            rows_alloc = std::max(r.Hbox * wbox, wbox - 1 + CONV_BM + (o.kh - 1) * wbox + o.kw - 1);
          } else {
            r.R = std::min(o.Ho, CONV_BM / wbox);
            r.Hbox = r.R + o.kh - 1;
            r.row_blocks = (o.Ho + r.R - 1) / r.R;
            rows_alloc = std::max(r.Hbox * wbox, CONV_BM + (o.kh - 1) * wbox + o.kw - 1);
          }
          r.a_stage_bytes = (rows_alloc * o.cin * 2 + 1023) / 1024 * 1024;
          if (o.cin < 64) {
            // 64- / 32-byte operand rows: a stage only needs 256-byte alignment (strip_smem_layout); take it when it buys ring depth
            const int tight = (rows_alloc * o.cin * 2 + 255) / 256 * 256;
            auto depth = [&](int bytes) {
              int st = (int)std::min<size_t>(8, (232448 - (strip_smem_layout(0, bytes, o.k_pad / 64, o.cout).total + 1024)) / bytes);
              return st > 2 ? st & ~1 : st;
            };
            if (depth(tight) > depth(r.a_stage_bytes)) r.a_stage_bytes = tight;
          }
          r.bn_tile = 0;                       // force a fresh weight map below
          // CTA pairs for the issue-bound layers (>= 8 K steps per tile, two accumulators per warp): cout / 2 weight rows per CTA
          r.pair = net->strip_pair && r.flat && B >= 2 && o.cout % 32 == 0 && o.kh * o.kw * o.cin / 16 >= 8 && (sms & ~1) >= 2;
          int rc = make_tmap_f16_2d(&r.tmap_w, net->d_weights + o.w_off, (uint64_t)o.cout, (uint64_t)o.k_pad, (uint64_t)o.k_pad * 2,
                                    (uint32_t)(r.pair ? o.cout / 2 : o.cout));
          if (rc != FIRE_OK) return rc;
          r.box_cols = o.cout % 64 == 0 ? 64 : (o.cout % 32 == 0 ? 32 : 16);
          const size_t fixed = strip_smem_layout(0, r.a_stage_bytes, o.k_pad / 64, o.cout).total + 1024;
          r.stages = (int)std::min<size_t>(8, (232448 - fixed) / r.a_stage_bytes);
          r.n_acc = std::min(STRIP_MAX_ACC, 512 / pow2_cols(o.cout));          // accumulators side by side in TMEM
          if (const char* e = getenv("FIRE_B200_STRIP_ACC")) r.n_acc = std::max(2, std::min(r.n_acc, atoi(e)));   // A/B experiments (power of two)
          while (r.n_acc > 2 && r.n_acc / 2 > r.stages) r.n_acc >>= 1;
          // Two MMA-issuing warps (conv_strip.cuh) each wait only for their own tiles' patches: tile t is issued by warp t % 2
          // and lives in stage t % stages, so the ring depth must be EVEN - then a stage always belongs to the same warp and no
          // warp ever skips a fill of a barrier it waits on (mbarrier waits are by phase parity; see block35_fused.cuh).
          if (o.kh * o.kw * o.cin / 16 >= 8 && r.n_acc >= 4 && net->strip_mma_warps == 2 && r.stages > 2) r.stages &= ~1;
          if (r.stages < 2) r.strip = false;
          else {
            r.smem = strip_smem_layout(r.stages, r.a_stage_bytes, o.k_pad / 64, o.cout).total + 1024;
            r.tmem_cols = pow2_cols(r.n_acc * o.cout);
            const BlobBuf& sb = net->bufs[o.src_buf];
            const BlobBuf& db = net->bufs[o.dst_buf];
            const __half* src = static_cast<const __half*>(buf_ptr(net, o.src_buf, B, in, ws, out_raw)) + o.src_coff;
            const __half* dst = static_cast<const __half*>(buf_ptr(net, o.dst_buf, B, in, ws, out_raw)) + o.dst_coff;
            rc = make_tmap_f16_nhwc(&r.tmap_a, src, (uint64_t)o.cin, (uint64_t)o.W, (uint64_t)o.H, (uint64_t)B, (uint64_t)sb.C,
                                    (uint64_t)buf_wp(sb), (uint32_t)o.cin, (uint32_t)r.Wbox, (uint32_t)r.Hbox);
            if (rc != FIRE_OK) return rc;
            if (r.flat)
              rc = make_tmap_f16_pos3d(&r.tmap_out, dst, (uint64_t)o.cout, (uint64_t)o.Ho * r.Wbox, (uint64_t)B, (uint64_t)db.C,
                                       (uint32_t)r.box_cols, CONV_BM);
            else
              rc = make_tmap_f16_nhwc(&r.tmap_out, dst, (uint64_t)o.cout, (uint64_t)o.Wo, (uint64_t)o.Ho, (uint64_t)B, (uint64_t)db.C,
                                      (uint64_t)o.Wo, (uint32_t)r.box_cols, (uint32_t)r.Wbox, (uint32_t)r.R);
            if (rc != FIRE_OK) return rc;
            r.n_tiles = 1; r.n_res = 0; r.n_issuers = 1;
            continue;
          }
        }
      }
      const bool residual = (o.flags & CF_RESIDUAL) != 0;
      int bn = pick_bn_tile(o.cout, r.m_tiles, o.k_pad / 64, sms, residual);
      if (const char* e = getenv("FIRE_B200_BN")) {            // tuning experiments: "op:bn,op:bn,..."
        const int me = (int)(&r - net->ops.data());
        for (const char* q = e; q && *q;) {
          int op = -1, v = 0;
          if (sscanf(q, "%d:%d", &op, &v) == 2 && op == me && v >= 16 && v <= 256 && v % 16 == 0 && o.cout % v == 0 && (!residual || v % 64 == 0)) bn = v;
          q = strchr(q, ',');
          if (q) ++q;
        }
      }
      // CTA pairs (cta_group::2): TMA-fed layers without residual whose pairs still fill the GPU
      const bool gpair = net->igemm_pair && !residual && !(o.flags & CF_OUT_F32) && bn % 32 == 0 && bn >= 64 &&
                         (long long)((r.m_tiles + 1) / 2) * (o.cout / bn) >= (sms / 2) - net->pair_slack && o.k_pad >= 256;
      if (bn != r.bn_tile || gpair != r.gpair) {
        int rc = make_tmap_f16_2d(&r.tmap_w, net->d_weights + o.w_off, (uint64_t)o.cout, (uint64_t)o.k_pad,
                                  (uint64_t)o.k_pad * 2, (uint32_t)(gpair ? bn / 2 : bn));
        if (rc != FIRE_OK) return rc;
        r.bn_tile = bn;
        r.gpair = gpair;
      }
      r.n_tiles = o.cout / bn;
      r.n_res = residual ? bn / 64 : 0;
      r.box_cols = bn % 64 == 0 ? 64 : (bn % 32 == 0 ? 32 : 16);
      const size_t stage = CONV_A_STAGE_BYTES + (size_t)(r.gpair ? bn / 2 : bn) * 128;
      const size_t fixed = conv_smem_layout(0, bn, o.cout, r.n_res, r.gpair).total + 1024;     // + alignment slack
      r.stages = (int)std::min<size_t>(r.gpair ? std::min(net->max_stages, 8) : net->max_stages, (232448 - fixed) / stage);
      if (r.stages < 2) return fail(FIRE_ERR_UNSUPPORTED, "conv tile %d does not fit shared memory", bn);
      // TMA issuing threads (1x1 layers): the ring depth is rounded down to a multiple of their number (see conv_igemm.cuh)
      r.n_issuers = 1;
      if ((r.tma_a || r.im2col) && net->n_issuers > 1) {
        if (r.stages > 4 && r.stages % 4 && r.stages % 3) --r.stages;        // 5 -> 4, 7 -> 6: a ring the issuers can share
        for (int j = std::min(net->n_issuers, CONV_MAX_ISSUERS); j >= 2; --j)
          if (r.stages % j == 0) { r.n_issuers = j; break; }
      }
      r.smem = conv_smem_layout(r.stages, bn, o.cout, r.n_res, r.gpair).total + 1024;
      r.tmem_cols = pow2_cols(2 * bn);
      const BlobBuf& sb = net->bufs[o.src_buf];
      const BlobBuf& db = net->bufs[o.dst_buf];
      const uint64_t M = (uint64_t)B * o.Ho * o.Wo;
      if (r.tma_a) {
        const __half* src = static_cast<const __half*>(buf_ptr(net, o.src_buf, B, in, ws, out_raw)) + o.src_coff;
        int rc = make_tmap_f16_2d(&r.tmap_a, src, M, (uint64_t)o.cin, (uint64_t)sb.C * 2, CONV_BM);
        if (rc != FIRE_OK) return rc;
      } else if (r.im2col) {
        if (buf_wp(sb) != sb.W) return fail(FIRE_ERR_UNSUPPORTED, "im2col conv over a pitched buffer");
        const __half* src = static_cast<const __half*>(buf_ptr(net, o.src_buf, B, in, ws, out_raw)) + o.src_coff;
        int rc = make_tmap_f16_im2col(&r.tmap_a, src, (uint64_t)o.cin, (uint64_t)o.W, (uint64_t)o.H, (uint64_t)B, (uint64_t)sb.C, o.kw, o.kh, o.pad_w, o.pad_h, o.stride);
        if (rc != FIRE_OK) return rc;
      }
      if (r.n_res) {
        const BlobBuf& rb = net->bufs[o.res_buf];
        const __half* res = static_cast<const __half*>(buf_ptr(net, o.res_buf, B, in, ws, out_raw)) + o.res_coff;
        int rc = make_tmap_f16_2d(&r.tmap_res, res, M, (uint64_t)o.cout, (uint64_t)rb.C * 2, CONV_BM);
        if (rc != FIRE_OK) return rc;
      }
      if (!(o.flags & CF_OUT_F32)) {
        const __half* dst = static_cast<const __half*>(buf_ptr(net, o.dst_buf, B, in, ws, out_raw)) + o.dst_coff;
        int rc = make_tmap_f16_2d_ex(&r.tmap_out, dst, M, (uint64_t)o.cout, (uint64_t)db.C * 2, (uint32_t)r.box_cols, CONV_BM);
        if (rc != FIRE_OK) return rc;
      }
    }
    if (net->f17.first_op >= 0) {
      Fused17& f = net->f17;
      const uint64_t M = (uint64_t)B * 64;
      for (int j = 0; j <= f.n_blocks; ++j) {
        const BlobOp& o = j < f.n_blocks ? net->ops[f.first_op + 4 * j].op : net->ops[f.first_op + 4 * (f.n_blocks - 1) + 3].op;
        const int buf = j < f.n_blocks ? o.src_buf : o.dst_buf;
        const __half* ptr = static_cast<const __half*>(buf_ptr(net, buf, B, in, ws, out_raw));
        f.prm.xptr[j] = ptr;
        int rc = make_tmap_f16_2d(&f.prm.xmap[j], ptr, M, (uint64_t)B17_C, (uint64_t)B17_C * 2, CONV_BM);
        if (rc != FIRE_OK) return rc;
      }
      f.prm.wstream = f.d_stream; f.prm.bias = f.d_bias; f.prm.n_blocks = f.n_blocks; f.prm.M_total = (int)M;
      f.prm.n_tiles = (int)((M + CONV_BM - 1) / CONV_BM); f.prm.trace = f.d_trace;
    }
    if (net->f35.first_op >= 0) {
      Fused35& f = net->f35;
      const uint64_t M = (uint64_t)B * B35_POS;
      for (int j = 0; j <= f.n_blocks; ++j) {
        const BlobOp& o = j < f.n_blocks ? net->ops[f.first_op + 4 * j].op : net->ops[f.first_op + 4 * (f.n_blocks - 1) + 3].op;
        const int buf = j < f.n_blocks ? o.src_buf : o.dst_buf;
        const __half* ptr = static_cast<const __half*>(buf_ptr(net, buf, B, in, ws, out_raw));
        f.prm.xptr[j] = ptr;
        int rc = make_tmap_f16_2d(&f.prm.xmap[j], ptr, M, (uint64_t)B35_C, (uint64_t)B35_C * 2, CONV_BM);
        if (rc != FIRE_OK) return rc;
      }
      f.prm.wstream = f.d_stream; f.prm.bias = f.d_bias; f.prm.n_blocks = f.n_blocks; f.prm.n_images = B; f.prm.trace = f.d_trace;
      if (f.balance && B > f.flags_cap) {
        cudaFree(f.d_flags);
        f.d_flags = nullptr; f.flags_cap = 0;
        if (cudaMalloc(&f.d_flags, sizeof(int) * (size_t)f.n_blocks * B) != cudaSuccess) { cudaGetLastError(); return fail(FIRE_ERR_CUDA, "block35 flags: out of device memory"); }
        FIRE_CUDA(cudaMemset(f.d_flags, 0, sizeof(int) * (size_t)f.n_blocks * B));
        f.flags_cap = B; f.epoch = 0;
      }
      f.prm.flags = f.d_flags; f.prm.balance = f.balance ? 1 : 0;
    }
    if (net->fpc.pool_op >= 0) {
      FusedPoolConv& f = net->fpc;
      const BlobOp& po = net->ops[f.pool_op].op;
      const BlobOp& co = net->ops[f.pool_op + 1].op;
      const BlobBuf& sb = net->bufs[po.src_buf];
      const void* xp = buf_ptr(net, po.src_buf, B, in, ws, out_raw);
      const void* yp = buf_ptr(net, co.dst_buf, B, in, ws, out_raw);
      int rc = make_tmap_f16_nhwc(&f.prm.in_map, xp, PC_CIN, PC_IN, PC_IN, (uint64_t)B, (uint64_t)sb.C, (uint64_t)buf_wp(sb), PC_CIN, PC_IN, PC_IN_ROWS);
      if (rc == FIRE_OK) rc = make_tmap_f16_2d(&f.prm.w_map, net->d_weights + co.w_off, PC_COUT, PC_CIN, PC_CIN * 2, PC_COUT);
      if (rc == FIRE_OK) rc = make_tmap_f16_pos3d(&f.prm.out_a, yp, PC_COUT, (uint64_t)PC_OUT * PC_OUT, (uint64_t)B, PC_COUT, 64, PC_POS);
      if (rc == FIRE_OK) rc = make_tmap_f16_pos3d(&f.prm.out_b, yp, PC_COUT, (uint64_t)PC_OUT * PC_OUT, (uint64_t)B, PC_COUT, 16, PC_POS);
      if (rc != FIRE_OK) return rc;
      f.prm.bias = f.d_bias; f.prm.n_images = B;
    }
    if (net->f8.first_op >= 0) {
      Fused8& f = net->f8;
      for (int j = 0; j < f.n_blocks; ++j) {
        const BlobOp& hop = net->ops[f.first_op + 4 * j].op;
        const BlobOp& uop = net->ops[f.first_op + 4 * j + 3].op;
        B8Params& q = f.prm[j];
        const BlobBuf& Xb = net->bufs[hop.dst_buf];
        const void* Xp = buf_ptr(net, hop.dst_buf, B, in, ws, out_raw);
        const void* xp = buf_ptr(net, hop.src_buf, B, in, ws, out_raw);
        const void* yp = buf_ptr(net, uop.dst_buf, B, in, ws, out_raw);
        int rc = make_tmap_f16_nhwc(&q.hmap_nat, Xp, (uint64_t)Xb.C, 3, 3, (uint64_t)B, (uint64_t)Xb.C, 3, 64, 3, 3, B8_IMGS);
        if (rc == FIRE_OK) rc = make_tmap_f16_nhwc(&q.hmap_ym, Xp, (uint64_t)Xb.C, 3, 3, (uint64_t)B, (uint64_t)Xb.C, 3, 64, 3, 1, B8_IMGS);
        if (rc == FIRE_OK) rc = make_tmap_f16_nhwc(&q.xmap_ym, xp, (uint64_t)B8_C, 3, 3, (uint64_t)B, (uint64_t)B8_C, 3, 64, 3, 1, B8_IMGS);
        if (rc == FIRE_OK) rc = make_tmap_f16_nhwc(&q.ymap_ym, yp, (uint64_t)B8_C, 3, 3, (uint64_t)B, (uint64_t)B8_C, 3, 64, 3, 1, B8_IMGS);
        if (rc != FIRE_OK) return rc;
        q.wmid = f.d_w + (size_t)j * B8_W_PER_BLOCK;
        q.wup = q.wmid + (size_t)B8_WMID_UNITS * B8_WMID_BYTES;
        q.bias = f.d_bias + (size_t)j * B8_BIAS_PER_BLOCK;
        q.n_groups = (B + B8_IMGS - 1) / B8_IMGS;
        q.relu = (uop.flags & CF_RELU) ? 1 : 0;
        q.b1a_coff = hop.dst_coff; q.b0_coff = hop.dst_coff + B8_MID;
        q.trace = f.d_trace ? f.d_trace + (size_t)j * 148 * B8_TRACE_SLOTS : nullptr;
        q.gap_out = nullptr; q.gap_ld = 0; q.n_images = B;
        if (j == f.n_blocks - 1 && f.gap_op >= 0) {
          const BlobOp& gop = net->ops[f.gap_op].op;
          q.gap_out = static_cast<__half*>(buf_ptr(net, gop.dst_buf, B, in, ws, out_raw));
          q.gap_ld = net->bufs[gop.dst_buf].C;
        }
      }
    }
    net->key_in = in; net->key_ws = ws; net->key_out = out_raw; net->key_B = B;
  }
  return FIRE_OK;
}

extern "C" {

int fire_facenet_forward(fire_net_t* net, const void* in_f16, int B, float* out_raw, float* out_l2, void* workspace,
                         size_t ws_bytes, fire_stream_t stream) {
  if (!net) return fail(FIRE_ERR_ARG, "fire_facenet_forward: NULL handle");
  int rc = use_device(net->device);
  if (rc != FIRE_OK) return rc;
  rc = prepare(net, in_f16, B, out_raw, workspace, ws_bytes);
  if (rc != FIRE_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool join_pending = false;
  for (size_t i = 0; i < net->ops.size(); ++i) {
    if (net->hoist_at[i] >= 0) {
      // everything launched so far (the producer of the pool's input included) -> fork; the pool runs beside the next convs
      FIRE_CUDA(cudaEventRecord(net->ev_fork, st));
      FIRE_CUDA(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
      rc = run_op(net, net->ops[net->hoist_at[i]], B, in_f16, workspace, out_raw, net->side, false);
      if (rc != FIRE_OK) return rc;
      FIRE_CUDA(cudaEventRecord(net->ev_join, net->side));
      join_pending = true;
    }
    if (net->hoisted[i]) {                              // its consumers come next: join here
      if (join_pending) { FIRE_CUDA(cudaStreamWaitEvent(st, net->ev_join, 0)); join_pending = false; }
      continue;
    }
    if (in_f17(net, i)) {
      if ((int)i == net->f17.first_op) { rc = run_f17(net, B, st, net->pdl); if (rc != FIRE_OK) return rc; }
      continue;
    }
    if (in_f35(net, i)) {
      if ((int)i == net->f35.first_op) { rc = run_f35(net, st, net->pdl); if (rc != FIRE_OK) return rc; }
      continue;
    }
    if (net->fpc.pool_op >= 0 && ((int)i == net->fpc.pool_op || (int)i == net->fpc.pool_op + 1)) {
      if ((int)i == net->fpc.pool_op + 1) { rc = run_pc(net, st, net->pdl); if (rc != FIRE_OK) return rc; }
      continue;
    }
    {
      const int role = f8_role(net, i);
      if (role == 2 || (int)i == net->f8.gap_op) continue;
      if (role == 1) { rc = run_f8(net, i, st, net->pdl); if (rc != FIRE_OK) return rc; continue; }
    }
    rc = run_op(net, net->ops[i], B, in_f16, workspace, out_raw, st, net->pdl);
    if (rc != FIRE_OK) return rc;
  }
  if (net->f8.d_trace) {
    FIRE_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> t((size_t)B8_MAX_BLOCKS * 148 * B8_TRACE_SLOTS);
    cudaMemcpy(t.data(), net->f8.d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    fprintf(stderr, "# block8_fused, ns since CTA 0 of block 1 entered.  entry | MMA warp: first operands, 1x3 issued, r2 ready, 3x1 issued, r3 ready, up issued |"
                    " epilogue warp 2: acc1, r2 written, acc2, r3 written, accU, staged, stores done   (CTA 0; then max over CTAs of entry / stores done)\n");
    const long long t0 = t[0];
    for (int j = 0; j < net->f8.n_blocks; ++j) {
      const long long* q = &t[(size_t)j * 148 * B8_TRACE_SLOTS];
      fprintf(stderr, "  blk %d | %7lld |", j + 1, q[0] - t0);
      for (int k = 1; k <= 6; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      fprintf(stderr, " |");
      for (int k = 8; k <= 14; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      long long emax = 0, smax = 0;
      const int ctas = std::min(148, net->f8.prm[j].n_groups * B8_NTILES);
      for (int c = 0; c < ctas; ++c) { emax = std::max(emax, q[c * B8_TRACE_SLOTS]); smax = std::max(smax, q[c * B8_TRACE_SLOTS + 14]); }
      fprintf(stderr, " | %7lld %7lld\n", emax - t0, smax - t0);
    }
  }
  if (net->f35.d_trace) {
    FIRE_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> t((size_t)148 * 16 * 24);
    cudaMemcpy(t.data(), net->f35.d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    fprintf(stderr, "# block35_fused CTA 0, ns since its first H.  MMA warp: H start, H issued, conv1 start, conv1 issued, conv2 start, conv2 issued, up start, up issued |"
                    " epilogue warp 2: accH[0], hready[2], acc1[0], conv1 epi done, acc2[0], aup_ready[2], up tiles 0..2 start, stores issued, y_done\n");
    const long long t0 = t[0];
    for (int j = 0; j < 2 * net->f35.n_blocks && j < 16; ++j) {
      const long long* q = &t[(size_t)j * 24];
      if (!q[0]) break;
      fprintf(stderr, "  blk %2d |", j);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      fprintf(stderr, " |");
      for (int k = 8; k <= 18; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      fprintf(stderr, " | up epilogue cycles (12 steps): fill+ldg %lld, acc wait %lld, first tmem ld %lld, chunks %lld, write-back %lld\n", q[19], q[20], q[21], q[22], q[23]);
    }
  }
  if (net->f17.d_trace) {
    FIRE_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> t((size_t)148 * B17_MAX_BLOCKS * B17_TRACE_SLOTS);
    cudaMemcpy(t.data(), net->f17.d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    fprintf(stderr, "# block17_fused CTA 0, ns since its first H.  MMA warp: H start, x landed, H issued, 1x7 start, up start, up issued, tile 2 wait, tile 2 go |"
                    " epilogue warp 2: accH, r1, r0, acc17, r2, acc71, r3, up tiles 0..3 start, stores issued, y_done\n");
    const long long t0 = t[0];
    for (int j = 0; j < net->f17.n_blocks; ++j) {
      const long long* q = &t[(size_t)j * B17_TRACE_SLOTS];
      fprintf(stderr, "  blk %d |", j);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      fprintf(stderr, " |");
      for (int k = 8; k <= 20; ++k) fprintf(stderr, " %7lld", q[k] - t0);
      fprintf(stderr, "\n");
    }
  }
  if (out_l2) {
    l2norm_kernel<<<(B + 7) / 8, 256, 0, st>>>(out_raw, out_l2, B, net->hdr.D);
    FIRE_LAUNCH_CHECK("l2norm_kernel");
    count_launch();
  }
  if (net->trace_all) {
    FIRE_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> t(net->ops.size() * 4096);
    cudaMemcpy(t.data(), net->d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    long long t0 = 0, prev_exit = 0;
    fprintf(stderr, "# op grid | first entry, setup, first full(max), last MMA commit(max), epilogue done(max), last exit [ns since op 0 entered] | span | gap to previous exit\n");
    for (size_t i = 0; i < net->ops.size(); ++i) {
      const OpRt& r = net->ops[i];
      if (r.op.kind != OP_CONV || in_f17(net, i) || in_f35(net, i) || f8_role(net, i) >= 1) continue;
      const int grid = (int)std::min<long long>(r.strip ? (long long)B * r.row_blocks : (long long)r.m_tiles * r.n_tiles, device_sm_count());
      const long long* q = &t[i * 4096];
      long long e0 = 1ll << 62, su = 0, ff = 0, mc = 0, ed = 0, ex = 0;
      for (int c = 0; c < grid; ++c) {
        e0 = std::min(e0, q[c * 8]); su = std::max(su, q[c * 8 + 1]); ff = std::max(ff, q[c * 8 + 3]);
        mc = std::max(mc, q[c * 8 + 4]); ed = std::max(ed, q[c * 8 + 6]); ex = std::max(ex, q[c * 8 + 7]);
      }
      if (t0 == 0) { t0 = e0; prev_exit = e0; }
      fprintf(stderr, "%3zu %3d | %8lld %8lld %8lld %8lld %8lld %8lld | %6lld | %6lld\n", i, grid, e0 - t0, su - t0, ff - t0, mc - t0,
              ed - t0, ex - t0, ex - e0, e0 - prev_exit);
      prev_exit = ex;
    }
  }
  return FIRE_OK;
}

int fire_facenet_profile(fire_net_t* net, const void* in_f16, int B, void* workspace, size_t ws_bytes, float* host_ms,
                         double* host_flops, int n_ops, fire_stream_t stream) {
  if (!net || !host_ms || n_ops < (int)net->ops.size()) return fail(FIRE_ERR_ARG, "fire_facenet_profile: bad arguments");
  { const int rc_dev = use_device(net->device); if (rc_dev != FIRE_OK) return rc_dev; }
  float* out_raw = nullptr;
  FIRE_CUDA(cudaMalloc(&out_raw, sizeof(float) * (size_t)B * net->hdr.D));
  int rc = prepare(net, in_f16, B, out_raw, workspace, ws_bytes);
  if (rc != FIRE_OK) { cudaFree(out_raw); return rc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev(net->ops.size() + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  cudaEventRecord(ev[0], st);
  for (size_t i = 0; i < net->ops.size(); ++i) {
    if (in_f17(net, i)) {                            // the fused chain is timed as a whole and reported on its first op
      if ((int)i == net->f17.first_op) rc = run_f17(net, B, st, false);
    } else if (in_f35(net, i)) {
      if ((int)i == net->f35.first_op) rc = run_f35(net, st, false);
    } else if (net->fpc.pool_op >= 0 && ((int)i == net->fpc.pool_op || (int)i == net->fpc.pool_op + 1)) {   // timed on the conv's slot
      if ((int)i == net->fpc.pool_op + 1) rc = run_pc(net, st, false);
    } else if (f8_role(net, i) >= 1 || (int)i == net->f8.gap_op) {   // the fused Block8 tail is timed on the 1x3 conv's slot
      if (f8_role(net, i) == 1) rc = run_f8(net, i, st, false);
    } else {
      rc = run_op(net, net->ops[i], B, in_f16, workspace, out_raw, st, false);   // no overlap: clean per-op times
    }
    if (rc != FIRE_OK) break;
    cudaEventRecord(ev[i + 1], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == FIRE_OK && e == cudaSuccess) {
    for (size_t i = 0; i < net->ops.size(); ++i) {
      cudaEventElapsedTime(&host_ms[i], ev[i], ev[i + 1]);
      if (host_flops) host_flops[i] = net->ops[i].flops_per_image * B;
    }
    if (host_flops && net->f17.first_op >= 0) {      // all the chain's FLOPs belong to the launch reported on its first op
      double sum = 0;
      for (int q = 0; q < 4 * net->f17.n_blocks; ++q) { sum += host_flops[net->f17.first_op + q]; host_flops[net->f17.first_op + q] = 0; }
      host_flops[net->f17.first_op] = sum;
    }
    if (host_flops && net->f35.first_op >= 0) {
      double sum = 0;
      for (int q = 0; q < 4 * net->f35.n_blocks; ++q) { sum += host_flops[net->f35.first_op + q]; host_flops[net->f35.first_op + q] = 0; }
      host_flops[net->f35.first_op] = sum;
    }
    if (host_flops && net->f8.first_op >= 0) {
      for (int j = 0; j < net->f8.n_blocks; ++j) {
        double* hf = host_flops + net->f8.first_op + 4 * j;
        hf[1] += hf[2] + hf[3]; hf[2] = 0; hf[3] = 0;
      }
    }
  }
  if (net->d_trace && rc == FIRE_OK && e == cudaSuccess) {
    std::vector<long long> t(8 * 512);
    cudaMemcpy(t.data(), net->d_trace, t.size() * 8, cudaMemcpyDeviceToHost);
    const OpRt& r = net->ops[net->trace_op];
    const int grid = (int)std::min<long long>(r.strip ? (long long)B * r.row_blocks : (long long)r.m_tiles * r.n_tiles, device_sm_count());
    long long t0 = t[0];
    for (int c = 0; c < grid; ++c) t0 = std::min(t0, t[c * 8]);
    const char* names[8] = {"entry", "setup done", "first TMA issued", "first full", "last MMA commit", "first acc_full", "epilogue done", "exit"};
    fprintf(stderr, "trace op %d: grid %d, tiles %d x %d, bn %d, nkb %d, stages %d, issuers %d, event ms %.4f\n", net->trace_op, grid,
            r.m_tiles, r.n_tiles, r.bn_tile, r.op.k_pad / 64, r.stages, r.n_issuers, host_ms[net->trace_op]);
    if (r.strip) {
      const char* ph[6] = {"MMA warp: wait acc_empty", "MMA warp: wait a_full", "MMA warp: issue + commit", "epilogue: wait out_empty", "epilogue: wait acc_full",
                           "epilogue: ld + cvt + sts + arrive"};
      const long long* q = &t[8 * 148];
      fprintf(stderr, "  cycles per tile (avg over CTAs), tiles per CTA %lld, accumulators %d, stages %d:\n", q[7], r.n_acc, r.stages);
      for (int k = 0; k < 6; ++k) {
        double avg = 0;
        for (int c = 0; c < grid; ++c) avg += (double)q[c * 8 + k] / (double)std::max<long long>(q[c * 8 + 7], 1);
        fprintf(stderr, "    %-36s %8.0f\n", ph[k], avg / grid);
      }
    }
    if (!r.strip) {
      const char* ph[3] = {"MMA warp: wait acc_empty (epilogue)", "MMA warp: wait operands (full)", "MMA warp: issue + commit (main K-blocks)"};
      const long long* q = &t[8 * 148];
      fprintf(stderr, "  cycles per tile (avg over CTAs), tiles per CTA %lld:\n", q[7]);
      for (int k = 0; k < 3; ++k) {
        double avg = 0;
        for (int c = 0; c < grid; ++c) avg += (double)q[c * 8 + k] / (double)std::max<long long>(q[c * 8 + 7], 1);
        fprintf(stderr, "    %-44s %8.0f\n", ph[k], avg / grid);
      }
    }
    for (int k = 0; k < 8; ++k) {
      long long mn = 1ll << 62, mx = 0; double avg = 0;
      for (int c = 0; c < grid; ++c) { long long v = t[c * 8 + k] - t0; mn = std::min(mn, v); mx = std::max(mx, v); avg += (double)v; }
      fprintf(stderr, "  %-18s min %7lld  avg %9.0f  max %7lld ns after the first CTA entered\n", names[k], mn, avg / grid, mx);
    }
  }
  for (auto& x : ev) cudaEventDestroy(x);
  cudaFree(out_raw);
  net->key_in = nullptr;   // out_raw was temporary: force descriptor rebuild next time
  if (rc != FIRE_OK) return rc;
  if (e != cudaSuccess) return fail(FIRE_ERR_CUDA, "fire_facenet_profile: %s", cudaGetErrorString(e));
  return FIRE_OK;
}

int fire_facenet_read_buffer(fire_net_t* net, int buf, int B, const void* in_f16, const void* workspace, void* host_out,
                             size_t bytes) {
  if (!net || buf < 0 || buf >= (int)net->bufs.size() || !host_out) return fail(FIRE_ERR_ARG, "fire_facenet_read_buffer: bad arguments");
  { const int rc_dev = use_device(net->device); if (rc_dev != FIRE_OK) return rc_dev; }
  const BlobBuf& b = net->bufs[buf];
  const size_t need = (size_t)B * b.H * buf_wp(b) * b.C * b.elt;
  if (bytes < need) return fail(FIRE_ERR_ARG, "fire_facenet_read_buffer: need %zu bytes", need);
  if (buf == net->hdr.out_buf) return fail(FIRE_ERR_ARG, "fire_facenet_read_buffer: the output buffer belongs to the caller");
  const void* src = buf_ptr(net, buf, B, in_f16, const_cast<void*>(workspace), nullptr);
  FIRE_CUDA(cudaMemcpy(host_out, src, need, cudaMemcpyDeviceToHost));
  return FIRE_OK;
}

int fire_ingest_f32(const float* in_nhwc3, int B, void* out_f16, fire_stream_t stream) {
  if (!in_nhwc3 || !out_f16 || B <= 0) return fail(FIRE_ERR_ARG, "fire_ingest_f32: bad arguments");
  const long long n_pos = (long long)B * 80 * 80;
  const int blocks = (int)std::min<long long>((n_pos + 255) / 256, 148 * 32);
  ingest_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_nhwc3, static_cast<__half*>(out_f16), n_pos);
  FIRE_LAUNCH_CHECK("ingest_f32_kernel");
  count_launch();
  return FIRE_OK;
}

}  // extern "C"
