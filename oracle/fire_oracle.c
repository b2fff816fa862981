/* ORACLE - test infrastructure only.  Never linked into, imported by or executed from the product
 * (libfire_b200.so / fire_b200).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * CPU restatement, in plain C, of the third-party arithmetic the reference's hot path calls:
 *
 *  (1) hnswlib 0.8.0 (requirements.txt:3; NOT vendored in /root/reference, restated from its
 *      published sources bruteforce.h / space_ip.h / python_bindings/bindings.cpp) as used by
 *      modules/hnsw_manager.py:20,127,137,147,237:
 *        - cosine space = inner-product space + normalisation of every added row and every query,
 *          x * (1.0f / (sqrtf(sum x^2) + 1e-30f))                       [bindings.cpp normalize_vector]
 *        - distance = 1.0f - <q, g>, fp32, the SIMD16 kernels' summation order for D % 16 == 0:
 *          16 independent lane sums, lanes added left to right           [space_ip.h SIMD16Ext AVX512]
 *        - BFIndex.knn_query = linear scan with a max-heap of (dist,label): the result is the k
 *          lexicographically smallest (dist, label) pairs, ascending     [bruteforce.h searchKnn]
 *      PARITY UNPINNED against hnswlib itself (the wheel cannot be installed here; the reference has
 *      no tests or golden vectors).  Pinned only by known-answer cases in tests/test_oracle_knn.py.
 *
 *  (2) OpenCV cv::resize(..., INTER_AREA) on 8UC3 (requirements.txt:9, call site
 *      modules/encoder.py:20), restated from imgproc/src/resize.cpp: equal size = copy, integer
 *      scale = resizeAreaFast (2x2: (a+b+c+d+2)>>2, else sum*(1.f/area) rounded half-even), fractional
 *      down-scale = resizeArea_ with float DecimateAlpha tables, any up-scaled axis = the 11-bit
 *      fixed-point linear path with INTER_AREA's own coefficient rule.
 *      PINNED: tests/test_oracle_preprocess.py checks it bit-for-bit against the live cv2 in this image.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off; no -ffast-math, so float order is as written).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------- */
/* (1) hnswlib cosine brute force                                                               */
/* ------------------------------------------------------------------------------------------- */

/* bindings.cpp: normalize_vector(float* data, float* norm_array) */
void fire_oracle_normalize(const float* in, float* out, size_t n, int D) {
  for (size_t r = 0; r < n; ++r) {
    const float* x = in + r * (size_t)D;
    float* y = out + r * (size_t)D;
    float norm = 0.0f;
    for (int i = 0; i < D; i++) norm += x[i] * x[i];
    norm = 1.0f / (sqrtf(norm) + 1e-30f);
    for (int i = 0; i < D; i++) y[i] = x[i] * norm;
  }
}

/* space_ip.h: InnerProductSIMD16Ext (AVX512 flavour: 16 lanes; D % 16 == 0), else the scalar loop */
static inline float inner_product(const float* a, const float* b, int D) {
  if (D % 16 == 0) {
    float lane[16];
    for (int j = 0; j < 16; j++) lane[j] = 0.0f;
    for (int i = 0; i < D; i += 16)
      for (int j = 0; j < 16; j++) lane[j] = lane[j] + a[i + j] * b[i + j];
    float sum = lane[0];
    for (int j = 1; j < 16; j++) sum += lane[j];
    return sum;
  }
  float res = 0.0f;
  for (int i = 0; i < D; i++) res += a[i] * b[i];
  return res;
}

static inline int pair_less(float da, uint64_t la, float db, uint64_t lb) {
  return da < db || (da == db && la < lb);
}

/* bruteforce.h searchKnn + bindings.cpp result ordering.  g and q are ALREADY normalised.
 * labels[r] is the label of row r (NULL = row index).  Returns 0, or -1 if k > n.
 * num_threads > 1 splits the QUERIES over pthreads (hnswlib's ParallelFor does the same). */
typedef struct {
  const float* g; const uint64_t* labels; size_t n; int D; const float* q; int q0, q1, k;
  uint64_t* out_labels; float* out_dist;
} bf_job_t;

static void* bf_worker(void* arg) {
  bf_job_t* J = (bf_job_t*)arg;
  const int k = J->k, D = J->D;
  for (int qi = J->q0; qi < J->q1; ++qi) {
    const float* qv = J->q + (size_t)qi * D;
    float* bd = J->out_dist + (size_t)qi * k;      /* kept sorted ascending by (dist, label) */
    uint64_t* bl = J->out_labels + (size_t)qi * k;
    int cnt = 0;
    for (size_t r = 0; r < J->n; ++r) {
      float dist = 1.0f - inner_product(qv, J->g + r * (size_t)D, D);
      uint64_t lab = J->labels ? J->labels[r] : (uint64_t)r;
      if (cnt == k && !pair_less(dist, lab, bd[k - 1], bl[k - 1])) continue;
      int j = cnt < k ? cnt : k - 1;
      while (j > 0 && pair_less(dist, lab, bd[j - 1], bl[j - 1])) { bd[j] = bd[j - 1]; bl[j] = bl[j - 1]; --j; }
      bd[j] = dist; bl[j] = lab;
      if (cnt < k) cnt++;
    }
  }
  return NULL;
}

int fire_oracle_bf_knn(const float* g, const uint64_t* labels, size_t n, int D, const float* q, int Q, int k,
                       uint64_t* out_labels, float* out_dist, int num_threads) {
  if (k < 1 || (size_t)k > n) return -1;
  if (num_threads < 1) num_threads = 1;
  if (num_threads > Q) num_threads = Q;
  if (num_threads > 256) num_threads = 256;
  bf_job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < num_threads; ++t) {
    bf_job_t J = {g, labels, n, D, q, (int)((long long)Q * t / num_threads), (int)((long long)Q * (t + 1) / num_threads), k,
                  out_labels, out_dist};
    jobs[t] = J;
  }
  for (int t = 1; t < num_threads; ++t) pthread_create(&th[t], NULL, bf_worker, &jobs[t]);
  bf_worker(&jobs[0]);
  for (int t = 1; t < num_threads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* Same result, different loop order: every thread walks the gallery ONCE for a block of 8 queries (each row is loaded
 * once per block instead of once per query), so that checking hundreds of queries against millions of rows is
 * compute-bound instead of DRAM-bound.  The arithmetic of every (query, row) pair - inner_product(), the pair order,
 * the insertion - is the code above, so the output is bit-identical to fire_oracle_bf_knn (tests/test_oracle_knn.py).
 * Used by the parity checks that compare the GPU against the oracle at BASELINE sizes. */
#define BF_QBLOCK 8
static void* bf_worker_blocked(void* arg) {
  bf_job_t* J = (bf_job_t*)arg;
  const int k = J->k, D = J->D;
  for (int qb = J->q0; qb < J->q1; qb += BF_QBLOCK) {
    const int nq = J->q1 - qb < BF_QBLOCK ? J->q1 - qb : BF_QBLOCK;
    int cnt[BF_QBLOCK] = {0};
    for (size_t r = 0; r < J->n; ++r) {
      const float* gv = J->g + r * (size_t)D;
      const uint64_t lab = J->labels ? J->labels[r] : (uint64_t)r;
      for (int b = 0; b < nq; ++b) {
        const int qi = qb + b;
        float* bd = J->out_dist + (size_t)qi * k;
        uint64_t* bl = J->out_labels + (size_t)qi * k;
        float dist = 1.0f - inner_product(J->q + (size_t)qi * D, gv, D);
        if (cnt[b] == k && !pair_less(dist, lab, bd[k - 1], bl[k - 1])) continue;
        int j = cnt[b] < k ? cnt[b] : k - 1;
        while (j > 0 && pair_less(dist, lab, bd[j - 1], bl[j - 1])) { bd[j] = bd[j - 1]; bl[j] = bl[j - 1]; --j; }
        bd[j] = dist; bl[j] = lab;
        if (cnt[b] < k) cnt[b]++;
      }
    }
  }
  return NULL;
}

int fire_oracle_bf_knn_blocked(const float* g, const uint64_t* labels, size_t n, int D, const float* q, int Q, int k,
                               uint64_t* out_labels, float* out_dist, int num_threads) {
  if (k < 1 || (size_t)k > n) return -1;
  if (num_threads < 1) num_threads = 1;
  if (num_threads > (Q + BF_QBLOCK - 1) / BF_QBLOCK) num_threads = (Q + BF_QBLOCK - 1) / BF_QBLOCK;
  if (num_threads > 256) num_threads = 256;
  bf_job_t jobs[256];
  pthread_t th[256];
  const int blocks = (Q + BF_QBLOCK - 1) / BF_QBLOCK;
  for (int t = 0; t < num_threads; ++t) {
    int q0 = (int)((long long)blocks * t / num_threads) * BF_QBLOCK, q1 = (int)((long long)blocks * (t + 1) / num_threads) * BF_QBLOCK;
    if (q1 > Q) q1 = Q;
    bf_job_t J = {g, labels, n, D, q, q0, q1, k, out_labels, out_dist};
    jobs[t] = J;
  }
  for (int t = 1; t < num_threads; ++t) pthread_create(&th[t], NULL, bf_worker_blocked, &jobs[t]);
  bf_worker_blocked(&jobs[0]);
  for (int t = 1; t < num_threads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* (2) cv::resize INTER_AREA, 8UC3                                                              */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int si, di; float alpha; } DecimateAlpha;

static int cv_round_d(double v) { return (int)lrint(v); }       /* round half to even (default FP mode) */
static int cv_round_f(float v) { return (int)lrintf(v); }
static int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }
static int cv_ceil_d(double v) { int i = (int)v; return i + (i < v); }
static uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static int compute_area_tab(int ssize, int dsize, int cn, double scale, DecimateAlpha* tab) {
  int k = 0;
  for (int dx = 0; dx < dsize; dx++) {
    double fsx1 = dx * scale;
    double fsx2 = fsx1 + scale;
    double cellWidth = scale < ssize - fsx1 ? scale : ssize - fsx1;
    int sx1 = cv_ceil_d(fsx1), sx2 = cv_floor_d(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    if (sx1 - fsx1 > 1e-3) {
      tab[k].di = dx * cn; tab[k].si = (sx1 - 1) * cn;
      tab[k++].alpha = (float)((sx1 - fsx1) / cellWidth);
    }
    for (int sx = sx1; sx < sx2; sx++) {
      tab[k].di = dx * cn; tab[k].si = sx * cn;
      tab[k++].alpha = (float)(1.0 / cellWidth);
    }
    if (fsx2 - sx2 > 1e-3) {
      double a = fsx2 - sx2; if (a > 1.) a = 1.; if (a > cellWidth) a = cellWidth;
      tab[k].di = dx * cn; tab[k].si = sx2 * cn;
      tab[k++].alpha = (float)(a / cellWidth);
    }
  }
  return k;
}

/* src: sh x sw x 3 uint8 with row stride sstride (bytes); dst: dh x dw x 3 contiguous.  Returns 0. */
int fire_oracle_resize_area_u8c3(const uint8_t* src, int sh, int sw, long sstride, uint8_t* dst, int dh, int dw) {
  const int cn = 3;
  if (sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0) return -1;
  if (sh == dh && sw == dw) {                       /* cv::resize: same size -> copyTo */
    for (int y = 0; y < sh; y++) memcpy(dst + (size_t)y * dw * cn, src + (size_t)y * sstride, (size_t)sw * cn);
    return 0;
  }
  double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  int iscale_x = cv_round_d(scale_x), iscale_y = cv_round_d(scale_y);
  int is_area_fast = fabs(scale_x - iscale_x) < 2.220446049250313e-16 && fabs(scale_y - iscale_y) < 2.220446049250313e-16;

  if (scale_x >= 1 && scale_y >= 1) {
    if (is_area_fast) {                             /* resizeAreaFast_Invoker */
      int area = iscale_x * iscale_y;
      float scale = 1.f / area;
      int dwidth1 = (sw / iscale_x);
      for (int dy = 0; dy < dh; dy++) {
        uint8_t* D = dst + (size_t)dy * dw * cn;
        int sy0 = dy * iscale_y;
        int w = sy0 + iscale_y <= sh ? dwidth1 : 0;
        if (sy0 >= sh) { memset(D, 0, (size_t)dw * cn); continue; }
        for (int dx = 0; dx < dw; dx++) {
          for (int c = 0; c < cn; c++) {
            if (dx < w) {
              int sum = 0;
              for (int yy = 0; yy < iscale_y; yy++)
                for (int xx = 0; xx < iscale_x; xx++)
                  sum += src[(size_t)(sy0 + yy) * sstride + (size_t)(dx * iscale_x + xx) * cn + c];
              if (iscale_x == 2 && iscale_y == 2) D[dx * cn + c] = (uint8_t)((sum + 2) >> 2);   /* SIMD 2x2 path */
              else D[dx * cn + c] = sat_u8(cv_round_f((float)sum * scale));
            } else {
              int sum = 0, count = 0, sx0 = dx * iscale_x;
              if (sx0 >= sw) { D[dx * cn + c] = 0; continue; }
              for (int yy = 0; yy < iscale_y; yy++) {
                if (sy0 + yy >= sh) break;
                for (int xx = 0; xx < iscale_x; xx++) {
                  if (sx0 + xx >= sw) break;
                  sum += src[(size_t)(sy0 + yy) * sstride + (size_t)(sx0 + xx) * cn + c]; count++;
                }
              }
              D[dx * cn + c] = sat_u8(cv_round_f((float)sum / count));
            }
          }
        }
      }
      return 0;
    }
    /* resizeArea_ with DecimateAlpha tables */
    DecimateAlpha* xtab = (DecimateAlpha*)malloc(sizeof(DecimateAlpha) * (size_t)(sw * 2 + 2));
    DecimateAlpha* ytab = (DecimateAlpha*)malloc(sizeof(DecimateAlpha) * (size_t)(sh * 2 + 2));
    float* buf = (float*)malloc(sizeof(float) * (size_t)dw * cn * 2);
    float* sum = buf + (size_t)dw * cn;
    int xtab_size = compute_area_tab(sw, dw, cn, scale_x, xtab);
    int ytab_size = compute_area_tab(sh, dh, 1, scale_y, ytab);
    int width = dw * cn;
    for (int dx = 0; dx < width; dx++) sum[dx] = 0.f;
    int prev_dy = ytab[0].di;
    for (int j = 0; j < ytab_size; j++) {
      float beta = ytab[j].alpha;
      int dy = ytab[j].di, sy = ytab[j].si;
      const uint8_t* S = src + (size_t)sy * sstride;
      for (int dx = 0; dx < width; dx++) buf[dx] = 0.f;
      for (int k = 0; k < xtab_size; k++) {
        int sxn = xtab[k].si, dxn = xtab[k].di;
        float alpha = xtab[k].alpha;
        float t0 = buf[dxn] + S[sxn] * alpha;
        float t1 = buf[dxn + 1] + S[sxn + 1] * alpha;
        float t2 = buf[dxn + 2] + S[sxn + 2] * alpha;
        buf[dxn] = t0; buf[dxn + 1] = t1; buf[dxn + 2] = t2;
      }
      if (dy != prev_dy) {
        uint8_t* D = dst + (size_t)prev_dy * width;
        for (int dx = 0; dx < width; dx++) { D[dx] = sat_u8(cv_round_f(sum[dx])); sum[dx] = beta * buf[dx]; }
        prev_dy = dy;
      } else {
        for (int dx = 0; dx < width; dx++) sum[dx] += beta * buf[dx];
      }
    }
    {
      uint8_t* D = dst + (size_t)prev_dy * width;
      for (int dx = 0; dx < width; dx++) D[dx] = sat_u8(cv_round_f(sum[dx]));
    }
    free(xtab); free(ytab); free(buf);
    return 0;
  }

  /* some axis is up-scaled: linear machinery, INTER_AREA coefficient rule, 11-bit fixed point */
  int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
  short* ialpha = (short*)malloc(sizeof(short) * (size_t)dw * 2);
  int* yofs = (int*)malloc(sizeof(int) * (size_t)dh);
  short* ibeta = (short*)malloc(sizeof(short) * (size_t)dh * 2);
  int xmax = dw;
  for (int dx = 0; dx < dw; dx++) {
    int sx = cv_floor_d(dx * scale_x);
    float fx = (float)((dx + 1) - (sx + 1) * inv_scale_x);
    fx = fx <= 0 ? 0.f : fx - cv_floor_d(fx);
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx + 1 >= sw) {
      if (xmax > dx) xmax = dx;
      if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    }
    xofs[dx] = sx;
    float c0 = 1.f - fx, c1 = fx;
    int a0 = cv_round_f(c0 * 2048), a1 = cv_round_f(c1 * 2048);
    ialpha[dx * 2] = (short)(a0 > 32767 ? 32767 : a0);
    ialpha[dx * 2 + 1] = (short)(a1 > 32767 ? 32767 : a1);
  }
  for (int dy = 0; dy < dh; dy++) {
    int sy = cv_floor_d(dy * scale_y);
    float fy = (float)((dy + 1) - (sy + 1) * inv_scale_y);
    fy = fy <= 0 ? 0.f : fy - cv_floor_d(fy);
    yofs[dy] = sy;
    float c0 = 1.f - fy, c1 = fy;
    ibeta[dy * 2] = (short)cv_round_f(c0 * 2048);
    ibeta[dy * 2 + 1] = (short)cv_round_f(c1 * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    int sy0 = yofs[dy];
    int r0 = sy0 < 0 ? 0 : (sy0 < sh ? sy0 : sh - 1);
    int r1 = sy0 + 1 < 0 ? 0 : (sy0 + 1 < sh ? sy0 + 1 : sh - 1);
    const uint8_t* S0 = src + (size_t)r0 * sstride;
    const uint8_t* S1 = src + (size_t)r1 * sstride;
    int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
    uint8_t* D = dst + (size_t)dy * dw * cn;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx] * cn;
      for (int c = 0; c < cn; c++) {
        int h0, h1;
        if (dx < xmax) {
          int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
          h0 = S0[sx + c] * a0 + S0[sx + cn + c] * a1;
          h1 = S1[sx + c] * a0 + S1[sx + cn + c] * a1;
        } else {
          h0 = S0[sx + c] * 2048;
          h1 = S1[sx + c] * 2048;
        }
        D[dx * cn + c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
      }
    }
  }
  free(xofs); free(ialpha); free(yofs); free(ibeta);
  return 0;
}

/* modules/face_recognition.py:412-420 crop rule + modules/encoder.py:19-27, for one box:
 * x,y,w,h each max(0,.) independently; numpy slicing clips the far edge; empty crop -> returns 1.
 * out_u8: 160x160x3 resized crop; out_f32: same /255.0f (may be NULL). */
int fire_oracle_crop_preprocess(const uint8_t* frame, int H, int W, long stride, int x, int y, int w, int h,
                                uint8_t* out_u8, float* out_f32) {
  if (x < 0) x = 0;
  if (y < 0) y = 0;
  if (w < 0) w = 0;
  if (h < 0) h = 0;
  int x1 = x + w > W ? W : x + w, y1 = y + h > H ? H : y + h;
  int x0 = x > W ? W : x, y0 = y > H ? H : y;
  int cw = x1 - x0, ch = y1 - y0;
  if (cw <= 0 || ch <= 0) return 1;
  int rc = fire_oracle_resize_area_u8c3(frame + (size_t)y0 * stride + (size_t)x0 * 3, ch, cw, stride, out_u8, 160, 160);
  if (rc) return rc;
  if (out_f32)
    for (int i = 0; i < 160 * 160 * 3; i++) out_f32[i] = (float)out_u8[i] / 255.0f;
  return 0;
}
