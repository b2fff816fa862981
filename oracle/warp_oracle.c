/* ORACLE - test infrastructure only.  Never linked into, imported by or executed from the product.
 *
 * CPU restatement of the enrol-side aligned crop of the reference (yunet_face_detector.py:135-165,
 * mediapipe_face_detector.py:155-182, retinaface_face_detector.py:298-326):
 *     M = cv2.getAffineTransform(pts1, pts2)            three float32 point pairs
 *     aligned = cv2.warpAffine(image, M, (160, 160))    INTER_LINEAR, BORDER_CONSTANT 0, uint8 x 3
 * restated from OpenCV's imgproc/src/imgwarp.cpp (getAffineTransform, invertAffineTransform inside warpAffine,
 * WarpAffineInvoker's 10-bit fixed-point coordinates with 5 interpolation bits, remapBilinear's 15-bit weight table
 * with its sum fix-up) and core/src/matrix_decomp.cpp (LU with partial pivoting, the solver getAffineTransform uses).
 * PINNED: tests/test_oracle_warp.py checks both functions bit for bit against the live cv2 in this image.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* cv::getAffineTransform: solve the 6 x 6 system with cv::solve(DECOMP_LU) = hal::LU64f */
int fire_oracle_get_affine(const float* src, const float* dst, double* M /* [6] row-major 2x3 */) {
  double a[36], b[6];
  memset(a, 0, sizeof(a));
  for (int i = 0; i < 3; ++i) {
    const int j = i * 12, k = i * 12 + 6;
    a[j] = a[k + 3] = src[2 * i];
    a[j + 1] = a[k + 4] = src[2 * i + 1];
    a[j + 2] = a[k + 5] = 1;
    b[i * 2] = dst[2 * i];
    b[i * 2 + 1] = dst[2 * i + 1];
  }
  const int m = 6;
  const double eps = 2.220446049250313e-16 * 100;        /* DBL_EPSILON * 100 */
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j)
      if (fabs(a[j * m + i]) > fabs(a[k * m + i])) k = j;
    if (fabs(a[k * m + i]) < eps) { memset(M, 0, 6 * sizeof(double)); return 0; }
    if (k != i) {
      for (int j = i; j < m; ++j) { double t = a[i * m + j]; a[i * m + j] = a[k * m + j]; a[k * m + j] = t; }
      double t = b[i]; b[i] = b[k]; b[k] = t;
    }
    const double d = -1 / a[i * m + i];
    for (int j = i + 1; j < m; ++j) {
      const double alpha = a[j * m + i] * d;
      for (int kk = i + 1; kk < m; ++kk) a[j * m + kk] += alpha * a[i * m + kk];
      b[j] += alpha * b[i];
    }
  }
  for (int i = m - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < m; ++k) s -= a[i * m + k] * b[k];
    b[i] = s / a[i * m + i];
  }
  memcpy(M, b, 6 * sizeof(double));
  return 1;
}

static int cv_round(double v) { return (int)lrint(v); }                 /* saturate_cast<int>(double): round half to even */
static short sat_s16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* remap's bilinear weight table BilinearTab_i[32*32][2][2] (INTER_REMAP_COEF_SCALE = 32768), built like
 * initInterTab2D does - including its sum fix-up, which for the 2 x 2 kernel looks at flat elements 3..6 counted from
 * the entry (its own (1,1) weight and the first three weights of the NEXT, not yet written, entry).  Net effect: a
 * short-fall (e.g. the saturated 32767 of an integer coordinate) is added to the (1,1) weight, an excess is dropped
 * unless the (1,1) weight is zero. */
static short g_flat[1024 * 4 + 8];
static int g_tab_ready = 0;
static void init_tab(void) {
  if (g_tab_ready) return;
  memset(g_flat, 0, sizeof(g_flat));
  for (int i = 0; i < 32; ++i) {
    const float xi = (float)i * (1.f / 32);
    const float ty[2] = {1.f - xi, xi};                     /* interpolateLinear */
    for (int j = 0; j < 32; ++j) {
      const float xj = (float)j * (1.f / 32);
      const float tx[2] = {1.f - xj, xj};
      short* it = g_flat + (i * 32 + j) * 4;
      int isum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const float v = ty[k1] * tx[k2];
          isum += it[k1 * 2 + k2] = sat_s16((int)lrintf(v * 32768.f));
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int Mk = 3, mk = 3;
        for (int idx = 3; idx <= 6; ++idx) {
          if (it[idx] < it[mk]) mk = idx;
          else if (it[idx] > it[Mk]) Mk = idx;
        }
        if (diff < 0) it[Mk] = (short)(it[Mk] - diff);
        else it[mk] = (short)(it[mk] - diff);
      }
    }
  }
  g_tab_ready = 1;
}

/* cv::warpAffine(src, M, (dw, dh)) with INTER_LINEAR, BORDER_CONSTANT (0): uint8, 3 channels */
int fire_oracle_warp_affine_u8c3(const uint8_t* src, int sh, int sw, long sstride, const double* M_fwd, uint8_t* dst, int dh, int dw) {
  init_tab();
  double M[6];
  memcpy(M, M_fwd, sizeof(M));
  {                                                        /* warpAffine inverts the forward matrix in double */
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D;
    M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5];
    const double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
  }
  const int AB_BITS = 10, AB_SCALE = 1 << 10, INTER_BITS = 5, round_delta = AB_SCALE / 32 / 2;
  for (int y = 0; y < dh; ++y) {
    const int X0 = cv_round((M[1] * y + M[2]) * AB_SCALE) + round_delta;
    const int Y0 = cv_round((M[4] * y + M[5]) * AB_SCALE) + round_delta;
    for (int x = 0; x < dw; ++x) {
      const int adelta = cv_round(M[0] * x * AB_SCALE), bdelta = cv_round(M[3] * x * AB_SCALE);
      const int X = (X0 + adelta) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS);
      const int sx = sat_s16(X >> INTER_BITS), sy = sat_s16(Y >> INTER_BITS);
      const short* w = g_flat + ((Y & 31) * 32 + (X & 31)) * 4;
      uint8_t* d = dst + ((size_t)y * dw + x) * 3;
      if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        const uint8_t* s = src + (size_t)sy * sstride + sx * 3;
        for (int c = 0; c < 3; ++c)
          d[c] = (uint8_t)((s[c] * w[0] + s[c + 3] * w[1] + s[sstride + c] * w[2] + s[sstride + c + 3] * w[3] + (1 << 14)) >> 15);
      } else if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
        d[0] = d[1] = d[2] = 0;                             /* entirely outside: the constant border value */
      } else {
        for (int c = 0; c < 3; ++c) {
          const int v0 = (sx >= 0 && sy >= 0 && sx < sw && sy < sh) ? src[(size_t)sy * sstride + sx * 3 + c] : 0;
          const int v1 = (sx + 1 >= 0 && sy >= 0 && sx + 1 < sw && sy < sh) ? src[(size_t)sy * sstride + (sx + 1) * 3 + c] : 0;
          const int v2 = (sx >= 0 && sy + 1 >= 0 && sx < sw && sy + 1 < sh) ? src[(size_t)(sy + 1) * sstride + sx * 3 + c] : 0;
          const int v3 = (sx + 1 >= 0 && sy + 1 >= 0 && sx + 1 < sw && sy + 1 < sh) ? src[(size_t)(sy + 1) * sstride + (sx + 1) * 3 + c] : 0;
          d[c] = (uint8_t)((v0 * w[0] + v1 * w[1] + v2 * w[2] + v3 * w[3] + (1 << 14)) >> 15);
        }
      }
    }
  }
  return 0;
}
