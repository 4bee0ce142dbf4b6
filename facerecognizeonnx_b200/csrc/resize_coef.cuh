// cv::resize(INTER_LINEAR, 8UC3) fixed-point coefficients and pixel evaluation, shared by the
// letterbox preprocess (reference src/face_detector.cpp:117) and the crop fallback / simple
// path (src/face_recognizer.cpp:123,170).  Recipe: SURVEY Appendix A.1, pinned against cv2 by
// tests/test_oracle_cv.py (oracle/cv_recipes.py is the CPU restatement).
#pragma once
#include <cstdint>

struct AxisCoef {
  int i0, i1;
  int w0, w1;
};

// cv::resize INTER_LINEAR coefficient for destination index d (see oracle/cv_recipes.py).
__device__ __forceinline__ AxisCoef axis_coef(int d, int n_dst, int n_src, bool horizontal) {
  const double scale = 1.0 / ((double)n_dst / (double)n_src);
  // separate IEEE multiply and subtract (no FMA contraction), as the x86 build of OpenCV does
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  AxisCoef c;
  if (horizontal) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    c.i0 = s;
    c.i1 = min(s + 1, n_src - 1);
  } else {
    c.i0 = min(max(s, 0), n_src - 1);
    c.i1 = min(max(s + 1, 0), n_src - 1);
  }
  c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}

__device__ __forceinline__ int resize_px(const uint8_t* r0, const uint8_t* r1, const AxisCoef& cx,
                                         const AxisCoef& cy, int ch) {
  const int h0 = r0[cx.i0 * 3 + ch] * cx.w0 + r0[cx.i1 * 3 + ch] * cx.w1;
  const int h1 = r1[cx.i0 * 3 + ch] * cx.w0 + r1[cx.i1 * 3 + ch] * cx.w1;
  return (((cy.w0 * (h0 >> 4)) >> 16) + ((cy.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

