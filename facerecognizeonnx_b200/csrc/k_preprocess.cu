// K1 det_preprocess: FaceDetector::preprocess (reference src/face_detector.cpp:92-137) as one
// pass: letterbox cv::resize (bit-exact 11-bit fixed point) + zero pad +
// BGR->RGB + (v-127.5)/128 + HWC->CHW + bf16 cast.  (v-127.5)/128 is exactly representable
// in bf16 for every byte value, so the bf16 output equals the reference's fp32 inputData.
//
// HBM-bound: 1,228,800 B read + 2,457,600 B written per 640x640 frame.
#include "common.h"
#include "resize_coef.cuh"

namespace {

constexpr int DET = FR_DET_SIZE;

__device__ __forceinline__ unsigned short norm_bf16_bits(int v) {
  // (v - 127.5) / 128, exact in bf16
  return __bfloat16_as_ushort(__float2bfloat16_rn(((float)v - 127.5f) * (1.0f / 128.0f)));
}

constexpr int PX_PER_THREAD = 8;
constexpr int THREADS_X = DET / PX_PER_THREAD;  // 80
constexpr int ROWS_PER_BLOCK = 4;

// grid: (DET / ROWS_PER_BLOCK, n_img); block: (80, 4).  Each thread produces 8 consecutive
// output pixels of one row for the three planes (3 x 16-byte stores).
__global__ void __launch_bounds__(THREADS_X* ROWS_PER_BLOCK)
det_preprocess_kernel(const ImgDesc* __restrict__ descs, __nv_bfloat16* __restrict__ out) {
  const ImgDesc d = descs[blockIdx.y];
  const int dy = blockIdx.x * ROWS_PER_BLOCK + threadIdx.y;
  const int dx0 = threadIdx.x * PX_PER_THREAD;
  unsigned short v[3][PX_PER_THREAD];
  const unsigned short padv = norm_bf16_bits(0);
  const bool identity = (d.new_w == d.cols) && (d.new_h == d.rows);
  if (dy >= d.new_h) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < PX_PER_THREAD; ++i) v[c][i] = padv;
  } else if (identity) {
    const uint8_t* row = d.ptr + (long long)dy * d.step;
    const int nvalid = min(PX_PER_THREAD, max(0, d.cols - dx0));
    if (nvalid == PX_PER_THREAD && ((reinterpret_cast<uintptr_t>(row + dx0 * 3) & 7) == 0)) {
      // 24 contiguous bytes, 8-byte aligned: three 64-bit loads
      const uint2* p = reinterpret_cast<const uint2*>(row + dx0 * 3);
      uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
      unsigned int wds[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
      for (int i = 0; i < PX_PER_THREAD; ++i)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int byte = i * 3 + ch;
          const int val = (wds[byte >> 2] >> ((byte & 3) * 8)) & 0xff;
          v[2 - ch][i] = norm_bf16_bits(val);  // BGR -> RGB plane
        }
    } else {
#pragma unroll
      for (int i = 0; i < PX_PER_THREAD; ++i) {
        if (i < nvalid) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) v[2 - ch][i] = norm_bf16_bits(row[(dx0 + i) * 3 + ch]);
        } else {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) v[ch][i] = padv;
        }
      }
    }
  } else {
    const AxisCoef cy = axis_coef(dy, d.new_h, d.rows, false);
    const uint8_t* r0 = d.ptr + (long long)cy.i0 * d.step;
    const uint8_t* r1 = d.ptr + (long long)cy.i1 * d.step;
#pragma unroll
    for (int i = 0; i < PX_PER_THREAD; ++i) {
      const int dx = dx0 + i;
      if (dx < d.new_w) {
        const AxisCoef cx = axis_coef(dx, d.new_w, d.cols, true);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[2 - ch][i] = norm_bf16_bits(resize_px(r0, r1, cx, cy, ch));
      } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch][i] = padv;
      }
    }
  }
  __nv_bfloat16* o = out + (size_t)blockIdx.y * 3 * DET * DET + (size_t)dy * DET + dx0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4 pk;
    pk.x = v[c][0] | ((unsigned)v[c][1] << 16);
    pk.y = v[c][2] | ((unsigned)v[c][3] << 16);
    pk.z = v[c][4] | ((unsigned)v[c][5] << 16);
    pk.w = v[c][6] | ((unsigned)v[c][7] << 16);
    *reinterpret_cast<uint4*>(o + (size_t)c * DET * DET) = pk;
  }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                   size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                   size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}

// Stand-alone cv::resize (u8, 3 channels) used by the crop fallback test hook.
__global__ void resize_u8_kernel(const uint8_t* __restrict__ src, int rows, int cols,
                                 long long step, int new_w, int new_h, uint8_t* __restrict__ dst) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= new_w || dy >= new_h) return;
  uint8_t* o = dst + ((size_t)dy * new_w + dx) * 3;
  if (new_w == cols && new_h == rows) {
    const uint8_t* p = src + (long long)dy * step + dx * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    return;
  }
  const AxisCoef cy = axis_coef(dy, new_h, rows, false);
  const AxisCoef cx = axis_coef(dx, new_w, cols, true);
  const uint8_t* r0 = src + (long long)cy.i0 * step;
  const uint8_t* r1 = src + (long long)cy.i1 * step;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) o[ch] = (uint8_t)resize_px(r0, r1, cx, cy, ch);
}

}  // namespace

int k_det_preprocess(fr_ctx* ctx, const ImgDesc* d_desc, int n_img, __nv_bfloat16* d_out_chw) {
  dim3 grid(DET / ROWS_PER_BLOCK, n_img), block(THREADS_X, ROWS_PER_BLOCK);
  det_preprocess_kernel<<<grid, block, 0, ctx->stream>>>(d_desc, d_out_chw);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_bf16_to_f32(fr_ctx* ctx, const __nv_bfloat16* in, float* out, size_t n) {
  bf16_to_f32_kernel<<<148 * 8, 256, 0, ctx->stream>>>(in, out, n);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_f32_to_bf16(fr_ctx* ctx, const float* in, __nv_bfloat16* out, size_t n) {
  f32_to_bf16_kernel<<<148 * 8, 256, 0, ctx->stream>>>(in, out, n);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_resize_u8(fr_ctx* ctx, const uint8_t* src, int rows, int cols, long long step, int new_w,
                int new_h, uint8_t* dst) {
  dim3 grid(ceil_div(new_w, 128), new_h), block(128);
  resize_u8_kernel<<<grid, block, 0, ctx->stream>>>(src, rows, cols, step, new_w, new_h, dst);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}
