// K5 align_warp: FaceRecognizer::alignFace (reference src/face_recognizer.cpp:93-133).
//
//  * align_estimate_kernel: cv::estimateAffinePartial2D(landmarks -> template) with all
//    defaults (:110-113) = RANSAC over 2-point similarity samples drawn from cv::RNG(-1)
//    (data-independent pair sequence), reprojection threshold 3 px, confidence 0.99, then the
//    least-squares similarity over the inliers (the fixed point of OpenCV's 10-iteration LM
//    refine).  Empty result -> the reference's crop fallback (:116-127).
//  * align_warp_kernel: cv::warpAffine INTER_LINEAR / BORDER_CONSTANT(0) in OpenCV's fixed
//    point (10-bit coordinates, 5-bit sub-pixel, 15-bit weights; Appendix A.2), or
//    cv::resize of (box & image) for the fallback.  Output: 112x112x3 BGR u8.
//
// This file is compiled with -fmad=false so the double/float expressions round exactly like
// the x86 build of OpenCV (and the numpy restatement in oracle/cv_recipes.py).
#include <cfloat>

#include "common.h"
#include "resize_coef.cuh"

namespace {

constexpr int REC = FR_REC_SIZE;

__constant__ float c_template[10] = {38.2946f, 51.6963f, 73.5318f, 51.5014f, 56.0252f,
                                     71.7366f, 41.5493f, 92.3655f, 70.7299f, 92.2041f};

struct CvRng {
  unsigned long long state;
  __device__ unsigned int next() {
    state = (unsigned long long)(unsigned int)state * 4164903690ull + (state >> 32);
    return (unsigned int)state;
  }
};

__global__ void align_estimate_kernel(const fr_face* __restrict__ faces,
                                      const int* __restrict__ face_img, int n_faces,
                                      const ImgDesc* __restrict__ descs, int n_img,
                                      AlignRec* __restrict__ recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_faces) return;
  const fr_face f = faces[i];
  const int img = face_img ? face_img[i] : 0;
  if ((unsigned)img >= (unsigned)n_img) {   // a device-side face->image map is caller data: never index with it unchecked
    AlignRec bad;
    bad.img = 0;
    bad.pad = 0;
    bad.mode = 2;
    bad.cx = bad.cy = bad.cw = bad.ch = 0;
    for (int k = 0; k < 6; ++k) bad.inv[k] = bad.fwd[k] = 0.0;
    recs[i] = bad;
    return;
  }
  float sx[5], sy[5], tx[5], ty[5];
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    sx[p] = f.lm[2 * p];
    sy[p] = f.lm[2 * p + 1];
    tx[p] = c_template[2 * p];
    ty[p] = c_template[2 * p + 1];
  }
  CvRng rng{0xffffffffffffffffull};
  int niters = 2000, max_good = 0;
  unsigned best = 0;
  for (int it = 0; it < niters; ++it) {
    const int i0 = rng.next() % 5u;
    int i1;
    do { i1 = rng.next() % 5u; } while (i1 == i0);
    // AffinePartial2DEstimatorCallback::runKernel (double)
    const double x1 = sx[i0], y1 = sy[i0], x2 = sx[i1], y2 = sy[i1];
    const double X1 = tx[i0], Y1 = ty[i0], X2 = tx[i1], Y2 = ty[i1];
    const double d = 1.0 / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
    const double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
    const double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
    const double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) -
                           (X1 * x2 - X2 * x1) * (x1 - x2));
    const double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) -
                           (Y1 * y2 - Y2 * y1) * (y1 - y2));
    // computeError: coefficients cast to float, fp32 error, <= thr^2
    const float F0 = (float)S0, F1 = (float)(-S1), F2 = (float)S2;
    const float F3 = (float)S1, F4 = (float)S0, F5 = (float)S3;
    unsigned mask = 0;
    int good = 0;
#pragma unroll
    for (int p = 0; p < 5; ++p) {
      const float a = F0 * sx[p] + F1 * sy[p] + F2 - tx[p];
      const float b = F3 * sx[p] + F4 * sy[p] + F5 - ty[p];
      const float e = a * a + b * b;
      if (e <= 9.0f) { mask |= 1u << p; ++good; }
    }
    if (good > max(max_good, 1)) {
      best = mask;
      max_good = good;
      // RANSACUpdateNumIters(0.99, (5-good)/5, 2, niters): cvRound(log(0.01)/log(1-(good/5)^2))
      // = 26, 10, 5 for good = 2, 3, 4 and 0 for good = 5 (denominator < DBL_MIN).
      const int upd = good == 2 ? 26 : good == 3 ? 10 : good == 4 ? 5 : 0;
      niters = min(niters, upd);
    }
  }
  AlignRec r;
  r.img = img;
  r.pad = 0;
  r.cx = r.cy = r.cw = r.ch = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) r.inv[k] = r.fwd[k] = 0.0;
  if (max_good == 0) {
    // fallback: (box & image) then resize (src/face_recognizer.cpp:116-127)
    const ImgDesc d = descs[img];
    const int x0 = max(f.x, 0), y0 = max(f.y, 0);
    const int x1 = min(f.x + f.w, d.cols), y1 = min(f.y + f.h, d.rows);
    if (x1 - x0 > 0 && y1 - y0 > 0) {
      r.mode = 1;
      r.cx = x0; r.cy = y0; r.cw = x1 - x0; r.ch = y1 - y0;
    } else {
      r.mode = 2;
    }
    recs[i] = r;
    return;
  }
  // least-squares similarity over the inliers (double)
  int n = 0;
  double mx = 0, my = 0, mX = 0, mY = 0;
#pragma unroll
  for (int p = 0; p < 5; ++p)
    if (best >> p & 1u) { mx += sx[p]; my += sy[p]; mX += tx[p]; mY += ty[p]; ++n; }
  mx /= n; my /= n; mX /= n; mY /= n;
  double den = 0, na = 0, nb = 0;
#pragma unroll
  for (int p = 0; p < 5; ++p)
    if (best >> p & 1u) {
      const double ax = sx[p] - mx, ay = sy[p] - my, bX = tx[p] - mX, bY = ty[p] - mY;
      den += ax * ax + ay * ay;
      na += ax * bX + ay * bY;
      nb += ax * bY - ay * bX;
    }
  const double a = na / den, b = nb / den;
  const double M0 = a, M1 = -b, M2 = mX - (a * mx - b * my);
  const double M3 = b, M4 = a, M5 = mY - (b * mx + a * my);
  r.fwd[0] = M0; r.fwd[1] = M1; r.fwd[2] = M2; r.fwd[3] = M3; r.fwd[4] = M4; r.fwd[5] = M5;
  // cv::warpAffine's inversion (imgwarp.cpp), double
  double D = M0 * M4 - M1 * M3;
  D = D != 0 ? 1.0 / D : 0;
  const double A11 = M4 * D, A22 = M0 * D, A12 = -M1 * D, A21 = -M3 * D;
  r.inv[0] = A11; r.inv[1] = A12; r.inv[2] = -A11 * M2 - A12 * M5;
  r.inv[3] = A21; r.inv[4] = A22; r.inv[5] = -A21 * M2 - A22 * M5;
  r.mode = 0;
  recs[i] = r;
}

__global__ void align_select_kernel(const fr_face* __restrict__ det, const int* __restrict__ n_det,
                                    int cap, const fr_face* __restrict__ pad, int n_img, int k,
                                    fr_face* __restrict__ sel, int* __restrict__ face_img,
                                    int* __restrict__ valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * k) return;
  const int img = i / k, j = i % k;
  fr_face f;
  int ok = 1;
  if (j < min(n_det[img], cap)) {
    f = det[(size_t)img * cap + j];
  } else if (pad) {
    f = pad[i];
  } else {
    f.x = f.y = f.w = f.h = 0;
    f.score = 0.f;
    for (int q = 0; q < 10; ++q) f.lm[q] = 0.f;
    ok = 0;
  }
  sel[i] = f;
  face_img[i] = img;
  valid[i] = ok;
}

__device__ __forceinline__ int tap(const uint8_t* base, long long step, int rows, int cols, int y,
                                   int x, int ch) {
  if ((unsigned)x >= (unsigned)cols || (unsigned)y >= (unsigned)rows) return 0;
  return base[(long long)y * step + x * 3 + ch];
}

// Six consecutive source bytes (two BGR pixels) starting at `p`, fetched with three aligned 32-bit
// loads and funnel shifts: lo = bytes 0..3, hi = bytes 4..7 (4 and 5 are used).
__device__ __forceinline__ void load6(const uint8_t* p, uint32_t& lo, uint32_t& hi) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}

// One output pixel of cv::warpAffine (INTER_LINEAR, BORDER_CONSTANT 0) given its fixed-point source
// coordinate: returns B | G << 8 | R << 16.
__device__ __forceinline__ uint32_t warp_px(const ImgDesc& d, int ix, int iy, int fx, int fy) {
  const int w00 = (32 - fx) * (32 - fy) * 32, w10 = fx * (32 - fy) * 32;
  const int w01 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  uint32_t out = 0;
  if (ix >= 1 && iy >= 0 && ix + 3 < d.cols && iy + 1 < d.rows) {
    // interior: the 2 x 2 taps are 2 rows x 6 contiguous bytes (the aligned 12-byte windows stay inside the row)
    const uint8_t* p0 = d.ptr + (long long)iy * d.step + ix * 3;
    uint32_t lo0, hi0, lo1, hi1;
    load6(p0, lo0, hi0);
    load6(p0 + d.step, lo1, hi1);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int t00 = (lo0 >> (8 * ch)) & 0xff, t01 = (lo1 >> (8 * ch)) & 0xff;
      const int t10 = ch == 0 ? (lo0 >> 24) : (hi0 >> (8 * (ch - 1))) & 0xff;
      const int t11 = ch == 0 ? (lo1 >> 24) : (hi1 >> (8 * (ch - 1))) & 0xff;
      const int acc = w00 * t00 + w10 * t10 + w01 * t01 + w11 * t11;
      out |= (uint32_t)((acc + (1 << 14)) >> 15) << (8 * ch);
    }
    return out;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int acc = w00 * tap(d.ptr, d.step, d.rows, d.cols, iy, ix, ch) +
                    w10 * tap(d.ptr, d.step, d.rows, d.cols, iy, ix + 1, ch) +
                    w01 * tap(d.ptr, d.step, d.rows, d.cols, iy + 1, ix, ch) +
                    w11 * tap(d.ptr, d.step, d.rows, d.cols, iy + 1, ix + 1, ch);
    out |= (uint32_t)((acc + (1 << 14)) >> 15) << (8 * ch);
  }
  return out;
}

// grid (REC / WARP_ROWS, n_faces), block (REC / 4, WARP_ROWS): one thread = 4 consecutive output
// pixels of a row = 12 output bytes, written as three 32-bit words (a warp's stores are contiguous).
constexpr int WARP_ROWS = 8;
__global__ void __launch_bounds__((REC / 4) * WARP_ROWS)
align_warp_kernel(const AlignRec* __restrict__ recs, const ImgDesc* __restrict__ descs,
                  uint8_t* __restrict__ crops, int* __restrict__ valid) {
  const int face = blockIdx.y;
  const AlignRec& r = recs[face];
  const int mode = r.mode;
  const int x0 = threadIdx.x * 4, y = blockIdx.x * WARP_ROWS + threadIdx.y;
  uint32_t* o = reinterpret_cast<uint32_t*>(crops + ((size_t)face * REC * REC + (size_t)y * REC + x0) * 3);
  uint32_t px[4] = {0u, 0u, 0u, 0u};   // B | G << 8 | R << 16 per pixel
  const bool dead = (valid && valid[face] == 0) || mode == 2;
  if (mode == 2 && x0 == 0 && y == 0 && valid && valid[face] != 0) valid[face] = 0;
  if (!dead) {
    const ImgDesc d = descs[r.img];
    if (mode == 1) {
      // crop (box & image) + cv::resize (src/face_recognizer.cpp:116-127)
      const uint8_t* src = d.ptr + (long long)r.cy * d.step + (long long)r.cx * 3;
      if (r.cw == REC && r.ch == REC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint8_t* p = src + (long long)y * d.step + (x0 + j) * 3;
          px[j] = p[0] | (p[1] << 8) | (p[2] << 16);
        }
      } else {
        const AxisCoef cy = axis_coef(y, REC, r.ch, false);
        const uint8_t* r0 = src + (long long)cy.i0 * d.step;
        const uint8_t* r1 = src + (long long)cy.i1 * d.step;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const AxisCoef cx = axis_coef(x0 + j, REC, r.cw, true);
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) px[j] |= (uint32_t)(resize_px(r0, r1, cx, cy, ch) & 0xff) << (8 * ch);
        }
      }
    } else {
      // mode 0: fixed-point affine warp (10-bit coordinates, 5-bit sub-pixel)
      const double A11 = r.inv[0], A12 = r.inv[1], b1 = r.inv[2];
      const double A21 = r.inv[3], A22 = r.inv[4], b2 = r.inv[5];
      const int X0 = __double2int_rn((A12 * (double)y + b1) * 1024.0) + 16;
      const int Y0 = __double2int_rn((A22 * (double)y + b2) * 1024.0) + 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        const int adelta = __double2int_rn(A11 * (double)x * 1024.0);
        const int bdelta = __double2int_rn(A21 * (double)x * 1024.0);
        const int X = (int)((unsigned)X0 + (unsigned)adelta) >> 5;
        const int Y = (int)((unsigned)Y0 + (unsigned)bdelta) >> 5;
        const int ix = min(max(X >> 5, -32768), 32767);
        const int iy = min(max(Y >> 5, -32768), 32767);
        px[j] = warp_px(d, ix, iy, X & 31, Y & 31);
      }
    }
  }
  // 4 pixels x 3 bytes -> 3 words
  o[0] = px[0] | (px[1] << 24);
  o[1] = (px[1] >> 8) | (px[2] << 16);
  o[2] = (px[2] >> 16) | (px[3] << 8);
}

}  // namespace

int k_align_estimate(fr_ctx* ctx, const fr_face* d_faces, const int* d_face_img, int n_faces,
                     const ImgDesc* d_desc, int n_img, AlignRec* d_rec) {
  if (n_faces <= 0) return FR_OK;
  align_estimate_kernel<<<ceil_div(n_faces, 64), 64, 0, ctx->stream>>>(d_faces, d_face_img,
                                                                       n_faces, d_desc, n_img, d_rec);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_align_select(fr_ctx* ctx, const fr_face* d_det, const int* d_n_det, int cap_per_img,
                   const fr_face* d_pad, int n_img, int k, fr_face* d_sel, int* d_face_img,
                   int* d_valid) {
  const int n = n_img * k;
  if (n <= 0) return FR_OK;
  align_select_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(d_det, d_n_det, cap_per_img,
                                                                 d_pad, n_img, k, d_sel,
                                                                 d_face_img, d_valid);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_align_warp(fr_ctx* ctx, const AlignRec* d_rec, int n_faces, const ImgDesc* d_desc,
                 uint8_t* d_crops, int* d_valid) {
  if (n_faces <= 0) return FR_OK;
  dim3 grid(REC / WARP_ROWS, n_faces), block(REC / 4, WARP_ROWS);
  align_warp_kernel<<<grid, block, 0, ctx->stream>>>(d_rec, d_desc, d_crops, d_valid);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}
