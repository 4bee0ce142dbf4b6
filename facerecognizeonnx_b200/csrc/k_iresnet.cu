// K6/K7: ArcFace w600k_r50 (IResNet-50) forward -- the replacement for
// `session_->Run` in FaceRecognizer::extractFeature (reference src/face_recognizer.cpp:270-283)
// plus FaceRecognizer::preprocess (:135-150, fused into the stem) and ::normalize (:306-318).
//
// All 3x3 / 1x1 convolutions and the FC run on the tcgen05 shift-GEMM kernel (tc_gemm.cuh):
//   * activations: NHWC bf16, padded flat layout (one shared zero halo row/column)
//   * the pre-conv BatchNorm of every IBasicBlock is folded exactly: its scale into conv1's
//     weights, its shift into a per-border-class bias table (the shift only reaches a pixel
//     through the taps that are inside the image)
//   * PReLU, bias and the residual add live in the GEMM epilogue; the stride-2 conv reads a
//     space-to-depth copy written by conv1's epilogue; the 1x1 stride-2 shortcut conv is a
//     tenth tap accumulating into the same TMEM tile
//   * BN2d -> flatten -> FC -> BN1d folds into one GEMM with K = 64 cells x 512 channels
// The 3->64 stem (0.3 % of the FLOPs, K = 27) is a SIMT kernel that also does the
// BGR->RGB / (v-127.5)/128 preprocess on the fly.
#include <cmath>
#include <cstring>

#include "common.h"
#include "tc_gemm.cuh"

namespace {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------ driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// 2-D bf16 tensor map [rows, cols] with row pitch `pitch_elems`, box = box_rows x 64, SW128.
bool tc_make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                    uint64_t pitch_elems, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D byte tensor map [rows, cols] (e4m3 operands), box = box_rows x 128 (128-byte rows), SW128.
bool tc_make_map_2d_u8(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t pitch_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D fp32 tensor map [rows, cols], box = box_rows x 32 (128-byte rows), SW128 (tf32 operands).
bool tc_make_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t pitch_elems, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 3-D fp32 tensor map over an NHWC activation with the image rows merged: dims (C, W, N*H),
// box (box_c, box_w, box_rows), no swizzle, out-of-range elements read as zero (conv padding).
bool tc_make_map_3d_f32(CUtensorMap* map, const void* base, uint64_t c, uint64_t w, uint64_t rows,
                        uint32_t box_c, uint32_t box_w, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {c, w, rows};
  cuuint64_t strides[2] = {c * 4, w * c * 4};
  cuuint32_t box[3] = {box_c, box_w, box_rows};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D tensor map of 8-byte elements [rows, cols] (dense), box = box_rows x box_cols, no swizzle,
// out-of-range elements read as zero: whole NHWC pixel rows of a 16-channel fp32 map.
bool tc_make_map_2d_u64(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                        uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 8};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct ConvLaunch {
  CUtensorMap a0, a1, b0, b1;
  CUtensorMap a_halo;      // halo mode: box = (a_rows / a_boxes) x 64
  CUtensorMap b_small;     // halo mode: the weights with a 64-row box (N-split tail items)
  bool has_b_small = false;
  // 2-CTA mode (halo_gemm2_kernel): the weights with BN/2-, BN/4- and BN/8-row boxes (each CTA of the pair
  // loads half of an item's weight rows; the smaller boxes serve the N-split tail items)
  CUtensorMap b_half, b_half2, b_half4;
  bool two_cta = false;
  // TMA-store epilogue (halo mode, OUT_STD, one N tile): the output tensor as [rows, cout], box 128 rows x 64 columns
  CUtensorMap out_map;
  bool tma_store = false;
  tc::Params p;
  int bn = 64;
  int rows_per_img = 1;  // Hp*Wp of the output geometry
  bool halo = false;
  int mt = 1;            // halo mode: M tiles per CTA iteration
  bool resb = false;     // halo mode: resident weights (Cin == 64)
};

static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Configure halo mode for a 3x3 stride-1 conv whose A operand is `base` ([rows, cin], pitch cin).
// the 256 / 512-channel 3x3 stride-1 layers run on CTA pairs (tc::halo_gemm2_kernel); FR_TC_2CTA=0 is the A/B switch
static bool two_cta_eligible(int bn, int cin) {
  (void)cin;
  static const int on = env_flag("FR_TC_2CTA", 1);
  return on && bn == 256;
}

static bool tc_setup_halo(ConvLaunch& L, const void* base, uint64_t rows, int cin, int Wp) {
  if (!env_flag("FR_TC_HALO", 1)) return true;   // default on (FR_TC_HALO=0 selects the per-tap loader)
  // M tiles per CTA iteration (they share one A block and every streamed weight tile).  Measured
  // per layer on B200: 2 wins only for the 128-channel 28x28 layers (114 -> 99 us); it loses at
  // 64 channels (resident weights, nothing to share) and at BN = 256 (no TMEM double buffer).
  int mt = env_flag("FR_TC_MT", 0);
  if (mt == 0) mt = (L.bn == 128 && Wp <= 29) ? 2 : 1;
  if (mt == 3) mt = L.bn <= 128 ? 2 : 1;
  L.mt = mt;
  L.resb = (cin == 64) && env_flag("FR_TC_RESB", 1);
  if (two_cta_eligible(L.bn, cin)) mt = 1;   // a CTA pair covers 256 rows with one tile per CTA
  int a_rows = (mt * tc::BM + 2 * Wp + 2 + 7) / 8 * 8;
  int boxes = 1;
  if (a_rows > 256) { boxes = 2; a_rows = (a_rows + 15) / 16 * 16; }
  L.p.a_rows = a_rows;
  L.p.a_boxes = boxes;
  L.p.base_off_mode = env_flag("FR_TC_BASEOFF", 0);
  L.halo = true;
  return tc_make_map_2d(&L.a_halo, base, rows, cin, cin, a_rows / boxes);
}

int tc_launch(fr_ctx* ctx, ConvLaunch& L, int m_rows) {
  L.p.m_rows = m_rows;
  {
    const unsigned long long per_img = (unsigned long long)std::max(L.p.Hp * L.p.Wp, 1);
    L.p.per_img_magic = ((1ull << 40) + per_img - 1) / per_img;
    const unsigned long long wp = (unsigned long long)std::max(L.p.Wp, 1);
    L.p.wp_magic = wp == 1 ? 0xffffffffu : (uint32_t)(((1ull << 32) + wp - 1) / wp);
  }
  L.p.num_m_tiles = ceil_div(m_rows, tc::BM);
  const int total = L.p.num_m_tiles * L.p.n_tiles_n * (L.p.k_splits > 1 ? L.p.k_splits : 1);
  if (total <= 0) return FR_OK;
  const int num_sms = ctx->num_sms;
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, tc::shift_gemm_kernel<64>, tc::Cfg<64>::SMEM_BYTES));
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, tc::shift_gemm_kernel<128>, tc::Cfg<128>::SMEM_BYTES));
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, tc::shift_gemm_kernel<256>, tc::Cfg<256>::SMEM_BYTES));
  const int grid = std::min(total, num_sms);
  if (L.halo && L.two_cta) {
    // cta_group::2: work item = 256 rows x BN columns per CTA pair (see tc::halo_gemm2_kernel)
    const int nclusters = num_sms / 2;
    const int supers = ceil_div(L.p.num_m_tiles, 2) * L.p.n_tiles_n;
    L.p.tail_split = 1;
    L.p.tail_first = supers;
    int items = supers;
    if (env_flag("FR_TC_TAILSPLIT", 1)) {
      const int tail = supers % nclusters;
      int split = 1;
      for (int s2 = 2; s2 <= 4 && L.bn / s2 >= 64; s2 *= 2)
        if (tail > 0 && tail * s2 <= nclusters) split = s2;
      if (split > 1) {
        L.p.tail_split = split;
        L.p.tail_first = supers - tail;
        items = L.p.tail_first + tail * split;
      }
    }
    const int cgrid = 2 * std::min(items, nclusters);
#define FR_HALO2_LAUNCH(BN_)                                                                                 \
  do {                                                                                                       \
    using C2 = tc::Halo2Cfg<BN_>;                                                                            \
    L.p.a_stages = std::min(env_flag("FR_TC_ASTAGES2", 2), C2::pick_a_stages(L.p.a_rows));                   \
    FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, tc::halo_gemm2_kernel<BN_>, 227 * 1024));                            \
    tc::halo_gemm2_kernel<BN_><<<cgrid, tc::CONV_THREADS, C2::smem_bytes(L.p.a_rows, L.p.a_stages),          \
                                ctx->stream>>>(L.a_halo, L.b_half, L.b_half2, L.b_half4, L.p);               \
  } while (0)
    FR_HALO2_LAUNCH(256);
#undef FR_HALO2_LAUNCH
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
    return FR_OK;
  }
  if (L.halo) {
    const int supers = ceil_div(L.p.num_m_tiles, L.mt) * L.p.n_tiles_n;
    // N-split of the last, partial wave (see tc::Params::tail_split)
    L.p.tail_split = 1;
    L.p.tail_first = supers;
    int items = supers;
    if (L.mt == 1 && !L.resb && L.has_b_small && env_flag("FR_TC_TAILSPLIT", 1)) {
      const int tail = supers % num_sms;
      int split = 1;
      for (int s2 = 2; s2 <= 4 && L.bn / s2 >= 64; s2 *= 2)
        if (tail > 0 && tail * s2 <= num_sms) split = s2;
      if (split > 1) {
        L.p.tail_split = split;
        L.p.tail_first = supers - tail;
        items = L.p.tail_first + tail * split;
      }
    }
    const int hgrid = std::min(items, num_sms);
#define FR_HALO_LAUNCH(BN_, MT_, RB_)                                                              \
  do {                                                                                             \
    using HC = tc::HaloCfg<BN_, MT_, RB_>;                                                         \
    FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, tc::halo_gemm_kernel<BN_, MT_, RB_>, 227 * 1024));         \
    /* staging tiles only where they fit beside two A blocks and the weights */                    \
    const bool ts = L.tma_store && BN_ <= 128 && HC::smem_bytes(L.p.a_rows, 2, true) <= 227 * 1024; \
    L.p.tma_store = ts ? 1 : 0;                                                                    \
    L.p.a_stages = std::min(env_flag("FR_TC_ASTAGES", 2), HC::pick_a_stages(L.p.a_rows, ts));      \
    tc::halo_gemm_kernel<BN_, MT_, RB_><<<hgrid, tc::CONV_THREADS,                                  \
        HC::smem_bytes(L.p.a_rows, L.p.a_stages, ts), ctx->stream>>>(                               \
        L.a_halo, L.b0, L.has_b_small ? L.b_small : L.b0, ts ? L.out_map : L.a_halo, L.p);          \
  } while (0)
    if (L.bn == 64 && L.resb) { if (L.mt == 2) FR_HALO_LAUNCH(64, 2, true); else FR_HALO_LAUNCH(64, 1, true); }
    else if (L.bn == 64) { if (L.mt == 2) FR_HALO_LAUNCH(64, 2, false); else FR_HALO_LAUNCH(64, 1, false); }
    else if (L.bn == 128) { if (L.mt == 2) FR_HALO_LAUNCH(128, 2, false); else FR_HALO_LAUNCH(128, 1, false); }
    else { if (L.mt == 2) FR_HALO_LAUNCH(256, 2, false); else FR_HALO_LAUNCH(256, 1, false); }
#undef FR_HALO_LAUNCH
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
    return FR_OK;
  }
  switch (L.bn) {
    case 64:
      tc::shift_gemm_kernel<64><<<grid, tc::CONV_THREADS, tc::Cfg<64>::SMEM_BYTES, ctx->stream>>>(
          L.a0, L.a1, L.b0, L.b1, L.p);
      break;
    case 128:
      tc::shift_gemm_kernel<128><<<grid, tc::CONV_THREADS, tc::Cfg<128>::SMEM_BYTES, ctx->stream>>>(
          L.a0, L.a1, L.b0, L.b1, L.p);
      break;
    default:
      tc::shift_gemm_kernel<256><<<grid, tc::CONV_THREADS, tc::Cfg<256>::SMEM_BYTES, ctx->stream>>>(
          L.a0, L.a1, L.b0, L.b1, L.p);
      break;
  }
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

namespace {

constexpr int REC = FR_REC_SIZE;

struct Act {          // one activation tensor in padded flat NHWC layout
  bf16* p = nullptr;
  int H = 0, W = 0, C = 0;  // valid extent (cells for s2d) and channels per row
  int Hp = 0, Wp = 0;
  size_t rows(int n) const { return (size_t)n * Hp * Wp; }
  size_t bytes(int n) const { return rows(n) * C * 2; }
};

struct BlockW {       // device weights of one IBasicBlock
  bf16* w1 = nullptr;     // [9][planes][cin]   (bn1 scale folded)
  float* b1 = nullptr;    // [9 classes][planes]
  float* prelu = nullptr; // [planes]
  std::vector<float> prelu_h;   // host copy (goes into the kernel parameters)
  bf16* w2 = nullptr;     // [9][planes][planes]
  bf16* wds = nullptr;    // [planes][cin] or null
  float* b2 = nullptr;    // [planes] (conv2 bias + ds bias)
  int cin = 0, planes = 0, stride = 1;
};

}  // namespace

struct RecModel {
  // stem
  float* stem_w = nullptr;  // [27][64], k = (r*3+s)*3 + c (c in RGB order)
  float* stem_b = nullptr;
  uint2* stem_bfrag = nullptr;  // stem weights as m16n8k16 B fragments [kstep][ntile][hi/lo][lane]
  float* stem_prelu = nullptr;
  float stem_prelu_h[64] = {0};   // host copy: stem_mma_kernel reads the slopes through the kernel-parameter bank
  std::vector<BlockW> blocks;
  bf16* fc_w = nullptr;     // [512][64*512] (bn2, feat affine folded; zero at halo cells)
  float* fc_b = nullptr;
  std::vector<void*> allocs;
  int* err_flag = nullptr;

  // plan (per capacity)
  int cap = 0;
  std::vector<void*> plan_allocs;
  Act x0, x0e;
  struct BlockBufs { Act h, out, out_even; };
  std::vector<BlockBufs> bufs;
  std::vector<ConvLaunch> conv1, conv2;
  ConvLaunch fc;
  float* fc_out = nullptr;   // [cap,512] raw
  float* fc_part = nullptr;  // [FC_SPLITS][cap,512] split-K partial sums
  float* chw_stage = nullptr;
  size_t chw_stage_cap = 0;
};

namespace {

template <typename T> T* dev_upload(fr_ctx* ctx, RecModel* m, const std::vector<T>& h) {
  T* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  m->allocs.push_back(d);
  return d;
}

std::vector<bf16> to_bf16(const std::vector<float>& v) {
  std::vector<bf16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
  return o;
}

// ------------------------------------------------------------------- stem kernel
// One thread = 4 horizontally adjacent output pixels x 64 channels (four passes of 16).  Input
// either u8 BGR HWC crops (FaceRecognizer::preprocess fused: BGR->RGB, (v-127.5)/128) or fp32
// CHW RGB.  ncu on the 2-pixel version: bound by the shared-memory weight reads (short-scoreboard
// stalls, one LDS.128 per 8 FMAs); 4 pixels per thread give 16 FMAs per weight float4.
template <bool U8>
__global__ void __launch_bounds__(128)
stem_kernel(const void* __restrict__ in_, int n, const float* __restrict__ w,
            const float* __restrict__ b, const float* __restrict__ slope, bf16* __restrict__ x0,
            bf16* __restrict__ x0e) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ __align__(16) float sb[64];
  __shared__ __align__(16) float ss[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 64) { sb[threadIdx.x] = b[threadIdx.x]; ss[threadIdx.x] = slope[threadIdx.x]; }
  __syncthreads();
  constexpr int PX = 4;
  const int quads_per_img = REC * (REC / PX);
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * quads_per_img) return;
  const int img = (int)(gid / quads_per_img);
  const int rem = (int)(gid % quads_per_img);
  const int y = rem / (REC / PX);
  const int x = (rem % (REC / PX)) * PX;
  // 3 rows x 6 cols x 3 channels input window (RGB order)
  float win[3][PX + 2][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < PX + 2; ++s) {
      const int yy = y + r - 1, xx = x + s - 1;
      const bool ok = yy >= 0 && yy < REC && xx >= 0 && xx < REC;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (ok) {
          if (U8) {
            const uint8_t* p = reinterpret_cast<const uint8_t*>(in_) +
                               ((size_t)img * REC * REC + (size_t)yy * REC + xx) * 3;
            v = ((float)p[2 - c] - 127.5f) * (1.0f / 128.0f);
          } else {
            v = reinterpret_cast<const float*>(in_)[((size_t)img * 3 + c) * REC * REC +
                                                    (size_t)yy * REC + xx];
          }
        }
        win[r][s][c] = v;
      }
    }
  const int Wp = REC + 1, Hp = REC + 1, We = REC / 2 + 1, He = REC / 2 + 1;
  bf16* o0 = x0 + ((size_t)(img * Hp + y) * Wp + x) * 64;
  bf16* oe = (!(y & 1)) ? x0e + ((size_t)(img * He + (y >> 1)) * We + (x >> 1)) * 64 : nullptr;
#pragma unroll 1
  for (int cg = 0; cg < 64; cg += 16) {
    float a[PX][16];
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[j][i] = sb[cg + i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4* wp = reinterpret_cast<const float4*>(&sw[((r * 3 + s) * 3 + c) * 64 + cg]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 w4 = wp[i];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
              const float v = win[r][s + j][c];
              a[j][4 * i] = fmaf(v, w4.x, a[j][4 * i]);
              a[j][4 * i + 1] = fmaf(v, w4.y, a[j][4 * i + 1]);
              a[j][4 * i + 2] = fmaf(v, w4.z, a[j][4 * i + 2]);
              a[j][4 * i + 3] = fmaf(v, w4.w, a[j][4 * i + 3]);
            }
          }
        }
#pragma unroll
    for (int j = 0; j < PX; ++j) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float sl = ss[cg + i];
        a[j][i] = a[j][i] > 0.f ? a[j][i] : a[j][i] * sl;
      }
      uint4 pk[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        pk[i].x = tc::pack_bf16(a[j][8 * i], a[j][8 * i + 1]);
        pk[i].y = tc::pack_bf16(a[j][8 * i + 2], a[j][8 * i + 3]);
        pk[i].z = tc::pack_bf16(a[j][8 * i + 4], a[j][8 * i + 5]);
        pk[i].w = tc::pack_bf16(a[j][8 * i + 6], a[j][8 * i + 7]);
      }
      reinterpret_cast<uint4*>(o0 + j * 64 + cg)[0] = pk[0];
      reinterpret_cast<uint4*>(o0 + j * 64 + cg)[1] = pk[1];
      if (oe && !(j & 1)) {
        reinterpret_cast<uint4*>(oe + (j >> 1) * 64 + cg)[0] = pk[0];
        reinterpret_cast<uint4*>(oe + (j >> 1) * 64 + cg)[1] = pk[1];
      }
    }
  }
}


// ------------------------------------------------------------------- stem on warp-level MMA
// The u8 hot path (FaceRecognizer::preprocess fused).  The layer is a 1 GB write with 43 MFLOP per
// face: its bound is HBM, but the CUDA-core kernel above is instruction bound (8000 instructions
// per 4 pixels, 0.67 ms).  Here a warp computes 16 pixels x 64 channels per step as an
// m16n8k16 bf16 GEMM with K = 27 (padded to 32): the normalised inputs (2v-255)/256 are exact in
// bf16, the weights enter as bf16 hi + lo, accumulation is fp32, so the result equals the fp32
// kernel's up to summation order.  Block = STEM_WARPS output rows of one face; the input strip (rows + 2, zero
// border) is staged once in shared memory as bf16 [row][(x+1)*3 + c] (RGB), where the 9 values of a
// filter row are contiguous.  The C fragments go through a padded per-warp tile so that every
// store instruction writes 64 contiguous bytes per pixel (whole sectors: lane-per-pixel 16-byte
// stores reach 1.4 TB/s, sector-complete ones > 6 TB/s, tests/dev_write_pattern.py).
constexpr int STEM_ROWP = (REC + 2) * 3;              // strip row pitch in elements
constexpr int STEM_WARPS = 8;                         // output rows per row group (one warp each)
constexpr int STEM_GROUPS = 2;                        // row groups per block (the next group's input is prefetched)
constexpr int STEM_STRIP = (STEM_WARPS + 3) * STEM_ROWP;   // rows + halo + a spare zero row for k = 27..31
constexpr int STEM_STAGE_PITCH = 144;                 // bytes per staged pixel (128 + 16: conflict-free)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

struct StemSlopes { float v[64]; };

__global__ void __launch_bounds__(STEM_WARPS * 32, 2)
stem_mma_kernel(const uint8_t* __restrict__ crops, const uint2* __restrict__ bfrag,
                const StemSlopes slope, bf16* __restrict__ x0, bf16* __restrict__ x0e) {
  __shared__ __align__(16) uint16_t strip[(STEM_STRIP + 7) / 8 * 8];
  __shared__ __align__(16) uint8_t stage[STEM_WARPS][16 * STEM_STAGE_PITCH];
  constexpr int NT = STEM_WARPS * 32;
  constexpr int ROW_WORDS = REC * 3 / 4;                             // 84 u32 per input row
  constexpr int FILL_WORDS = (STEM_WARPS + 2) * ROW_WORDS;
  constexpr int FILL_PER_THREAD = (FILL_WORDS + NT - 1) / NT;
  constexpr int BLOCKS_PER_IMG = REC / (STEM_WARPS * STEM_GROUPS);
  const int img = blockIdx.x / BLOCKS_PER_IMG;
  const int yblk = (blockIdx.x - img * BLOCKS_PER_IMG) * (STEM_WARPS * STEM_GROUPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint8_t* face = crops + (size_t)img * REC * REC * 3;
  // input words of the row group starting at output row yb (strip row i = input row yb - 1 + i)
  uint32_t wv[FILL_PER_THREAD];
  auto prefetch = [&](int yb) {
#pragma unroll
    for (int k = 0; k < FILL_PER_THREAD; ++k) {
      const int i = threadIdx.x + k * NT;
      const int row = i / ROW_WORDS, wd = i - row * ROW_WORDS, yy = yb - 1 + row;
      wv[k] = (i < FILL_WORDS && yy >= 0 && yy < REC)
                  ? __ldg(reinterpret_cast<const uint32_t*>(face + (size_t)yy * REC * 3) + wd) : 0u;
    }
  };
  prefetch(yblk);
  for (int i = threadIdx.x; i < (STEM_STRIP + 7) / 8; i += NT) reinterpret_cast<uint4*>(strip)[i] = make_uint4(0, 0, 0, 0);
  // The weight fragments [kstep][ntile][hi/lo] of this lane stay in registers for the whole block (64 registers):
  // read from shared memory per tile they were half of the kernel's shared-memory wavefronts, and the kernel
  // is bound by that pipe (ncu: L1 / shared 93 % busy, 131 wavefronts per 16-pixel tile, 64 of them these).
  uint2 bfr[2][8][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) bfr[ks][j][h] = __ldg(bfrag + ((ks * 8 + j) * 2 + h) * 32 + lane);
  // per-thread constants: strip offsets of its 8 k indices
  int koff[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int k = (q >> 2) * 16 + ((q >> 1) & 1) * 8 + 2 * t + (q & 1);
    koff[q] = (k / 9) * STEM_ROWP + (k % 9);
  }
  constexpr int Wp = REC + 1, Hp = REC + 1, We = REC / 2 + 1, He = REC / 2 + 1;
  uint8_t* st = stage[warp];
#pragma unroll 1
  for (int gr = 0; gr < STEM_GROUPS; ++gr) {
    const int y0 = yblk + gr * STEM_WARPS;
    __syncthreads();                                   // zero fill done / previous group's readers done
    // BGR bytes -> RGB bf16 of (v - 127.5) / 128 (exact in bf16); rows outside the face are zero
#pragma unroll
    for (int k = 0; k < FILL_PER_THREAD; ++k) {
      const int i = threadIdx.x + k * NT;
      if (i < FILL_WORDS) {
        const int row = i / ROW_WORDS, wd = i - row * ROW_WORDS, yy = y0 - 1 + row;
        const bool in = yy >= 0 && yy < REC;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = wd * 4 + q, px = e / 3, cb = e - px * 3;
          const float f = in ? ((float)((wv[k] >> (8 * q)) & 0xffu) - 127.5f) * (1.0f / 128.0f) : 0.f;
          strip[row * STEM_ROWP + (px + 1) * 3 + (2 - cb)] = __bfloat16_as_ushort(__float2bfloat16_rn(f));
        }
      }
    }
    __syncthreads();
    if (gr + 1 < STEM_GROUPS) prefetch(y0 + STEM_WARPS);   // in flight while this group computes
    const int y = y0 + warp;
    bf16* orow = x0 + (size_t)(img * Hp + y) * Wp * 64;
    bf16* erow = (!(y & 1)) ? x0e + (size_t)(img * He + (y >> 1)) * We * 64 : nullptr;
#pragma unroll 1
    for (int m = 0; m < REC / 16; ++m) {
      const uint16_t* s0 = strip + warp * STEM_ROWP + (16 * m + g) * 3;
      uint32_t a[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {                     // a0:(g,k lo) a1:(g+8,k lo) a2:(g,k hi) a3:(g+8,k hi)
          const uint16_t* sp = s0 + (q & 1) * 24;
          const int ko = ks * 4 + (q >> 1) * 2;
          a[ks][q] = (uint32_t)sp[koff[ko]] | ((uint32_t)sp[koff[ko + 1]] << 16);
        }
      }
      if (t == 1) {                                       // k = 27 carries the bias: A = 1.0
        a[1][2] = (a[1][2] & 0xffffu) | 0x3f800000u;
        a[1][3] = (a[1][3] & 0xffffu) | 0x3f800000u;
      }
      __syncwarp();                                       // the previous tile's staged pixels have been read
      // two halves of 32 channels: 16 accumulators live at a time (same accumulation order as before:
      // per n tile k step 0 hi, lo, k step 1 hi, lo)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float acc[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { acc[jj][0] = 0.f; acc[jj][1] = 0.f; acc[jj][2] = 0.f; acc[jj][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            mma_bf16_16816(acc[jj], a[ks], bfr[ks][4 * half + jj][0]);
            mma_bf16_16816(acc[jj], a[ks], bfr[ks][4 * half + jj][1]);
          }
        // PReLU, bf16, stage: pixel r of the tile at st + r*144, channel pair (8j + 2t) at byte 16j + 4t
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * half + jj;
          const float2 sl = *reinterpret_cast<const float2*>(&slope.v[8 * j + 2 * t]);   // constant bank (LDC)
          float v0 = acc[jj][0], v1 = acc[jj][1], v2 = acc[jj][2], v3 = acc[jj][3];
          v0 = v0 > 0.f ? v0 : v0 * sl.x; v1 = v1 > 0.f ? v1 : v1 * sl.y;
          v2 = v2 > 0.f ? v2 : v2 * sl.x; v3 = v3 > 0.f ? v3 : v3 * sl.y;
          *reinterpret_cast<uint32_t*>(st + g * STEM_STAGE_PITCH + 16 * j + 4 * t) = tc::pack_bf16(v0, v1);
          *reinterpret_cast<uint32_t*>(st + (g + 8) * STEM_STAGE_PITCH + 16 * j + 4 * t) = tc::pack_bf16(v2, v3);
        }
      }
      __syncwarp();
      // a quarter warp moves one pixel's 128 bytes: conflict-free reads, whole-line stores
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int r = pass * 4 + (lane >> 3), x = 16 * m + r;
        const uint4 v = *reinterpret_cast<const uint4*>(st + r * STEM_STAGE_PITCH + 16 * (lane & 7));
        *reinterpret_cast<uint4*>(orow + (size_t)x * 64 + 8 * (lane & 7)) = v;
        if (erow && !(x & 1)) *reinterpret_cast<uint4*>(erow + (size_t)(x >> 1) * 64 + 8 * (lane & 7)) = v;
      }
    }
  }
}

constexpr int FC_SPLITS = 16;

// sum of the split-K partial products of the FC (bias was added by split 0)
__global__ void fc_reduce_kernel(const float* __restrict__ part, size_t n_elems, size_t stride,
                                 float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < FC_SPLITS; ++k) s += part[(size_t)k * stride + i];
  out[i] = s;
}

// R4: FaceRecognizer::normalize (src/face_recognizer.cpp:306-318): one warp per row.
__global__ void l2_normalize_kernel(const float* __restrict__ in, int n, int dim,
                                    float* __restrict__ out, const int* __restrict__ valid) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* r = in + (size_t)row * dim;
  float* o = out + (size_t)row * dim;
  if (valid && valid[row] == 0) {
    for (int i = lane; i < dim; i += 32) o[i] = 0.f;
    return;
  }
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s = fmaf(r[i], r[i], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float norm = sqrtf(s);
  for (int i = lane; i < dim; i += 32) o[i] = norm > 0.f ? __fdiv_rn(r[i], norm) : r[i];
}

// K8 batched: out[i] = (dot(a_i, b_i) + 1) / 2 (src/face_recognizer.cpp:326-333).
__global__ void compare_batch_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     int n, int dim, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s = fmaf(a[(size_t)row * dim + i], b[(size_t)row * dim + i], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[row] = (s + 1.0f) / 2.0f;
}

// --------------------------------------------------------------- weight packing
// conv OIHW fp32 -> [tap][co][ci] bf16 with optional per-ci scale
std::vector<bf16> pack_conv3(const fr_tensor& w, const std::vector<float>* scale) {
  const int co = (int)w.dims[0], ci = (int)w.dims[1], k = (int)w.dims[2];
  std::vector<bf16> o((size_t)k * k * co * ci);
  for (int r = 0; r < k; ++r)
    for (int s = 0; s < k; ++s)
      for (int a = 0; a < co; ++a)
        for (int c = 0; c < ci; ++c) {
          float v = w.data[(((size_t)a * ci + c) * k + r) * k + s];
          if (scale) v *= (*scale)[c];
          o[(((size_t)(r * k + s)) * co + a) * ci + c] = __float2bfloat16_rn(v);
        }
  return o;
}

// bias table [9][co]: b[co] + sum over taps valid in the border class of W[co,ci,r,s]*shift[ci]
std::vector<float> border_bias(const fr_tensor& w, const std::vector<float>& b,
                               const std::vector<float>& shift) {
  const int co = (int)w.dims[0], ci = (int)w.dims[1];
  std::vector<double> tapsum((size_t)9 * co, 0.0);
  for (int a = 0; a < co; ++a)
    for (int c = 0; c < ci; ++c)
      for (int t = 0; t < 9; ++t) tapsum[(size_t)t * co + a] += (double)w.data[((size_t)a * ci + c) * 9 + t] * shift[c];
  std::vector<float> o((size_t)9 * co);
  for (int vc = 0; vc < 3; ++vc)
    for (int hc = 0; hc < 3; ++hc)
      for (int a = 0; a < co; ++a) {
        double s = b[a];
        for (int r = 0; r < 3; ++r) {
          if ((vc == 0 && r == 0) || (vc == 2 && r == 2)) continue;
          for (int q = 0; q < 3; ++q) {
            if ((hc == 0 && q == 0) || (hc == 2 && q == 2)) continue;
            s += tapsum[(size_t)(r * 3 + q) * co + a];
          }
        }
        o[(size_t)(vc * 3 + hc) * co + a] = (float)s;
      }
  return o;
}

void fill_taps_s1(tc::Params& p, int Wp, int cout, int cin) {
  p.num_taps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      tc::Tap& t = p.taps[r * 3 + s];
      t.a_src = 0; t.b_src = 0;
      t.a_row_shift = (r - 1) * Wp + (s - 1);
      t.a_col = 0;
      t.b_row = (r * 3 + s) * cout;
      t.nkb = cin / 64;
    }
}

// stride-2 3x3 conv over a space-to-depth input (4 phase blocks of `c` channels per cell)
void fill_taps_s2(tc::Params& p, int Wp2, int cout, int c) {
  p.num_taps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      tc::Tap& t = p.taps[r * 3 + s];
      t.a_src = 0; t.b_src = 0;
      const int dr = (r == 0) ? -1 : 0, dc = (s == 0) ? -1 : 0;
      const int pr = (r == 1) ? 0 : 1, pc = (s == 1) ? 0 : 1;
      t.a_row_shift = dr * Wp2 + dc;
      t.a_col = (pr * 2 + pc) * c;
      t.b_row = (r * 3 + s) * cout;
      t.nkb = c / 64;
    }
}

int pick_bn(int cout) { return cout >= 256 ? 256 : cout; }

}  // namespace

// ------------------------------------------------------------------ model create
int rec_model_create(fr_ctx* ctx, const fr_weights* w) {
  if (!w || w->model != FR_MODEL_REC) return fr_fail(ctx, FR_ERR_MODEL, "rec weights missing");
  if (!get_encode_fn()) return fr_fail(ctx, FR_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  std::unique_ptr<RecModel> m(new RecModel());
  {
    // stem: OIHW [64,3,3,3] -> [k=(r*3+s)*3+c][64]
    const fr_tensor& sw = w->at("stem.w");
    std::vector<float> pw(27 * 64);
    for (int co = 0; co < 64; ++co)
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s)
            pw[((r * 3 + s) * 3 + c) * 64 + co] = sw.data[((co * 3 + c) * 3 + r) * 3 + s];
    m->stem_w = dev_upload(ctx, m.get(), pw);
    // B fragments of mma.m16n8k16 (lane = 4g + t): b0 = W[k0 + 2t, +1][n = 8j + g], b1 = W[k0 + 2t + 8, +9][n]
    std::vector<uint2> frag(2 * 8 * 2 * 32);
    const std::vector<float>& sbias = w->at("stem.b").data;
    auto split = [&](int k, int n, int which) -> uint32_t {
      const float v = k < 27 ? pw[(size_t)k * 64 + n] : (k == 27 ? sbias[n] : 0.f);   // k = 27: bias row (A = 1)
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 r = which == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
      uint16_t u;
      memcpy(&u, &r, 2);
      return u;
    };
    for (int ks = 0; ks < 2; ++ks)
      for (int j = 0; j < 8; ++j)
        for (int h = 0; h < 2; ++h)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, k0 = ks * 16 + 2 * t, n = 8 * j + g;
            uint2 f;
            f.x = split(k0, n, h) | (split(k0 + 1, n, h) << 16);
            f.y = split(k0 + 8, n, h) | (split(k0 + 9, n, h) << 16);
            frag[((ks * 8 + j) * 2 + h) * 32 + lane] = f;
          }
    m->stem_bfrag = dev_upload(ctx, m.get(), frag);
    m->stem_b = dev_upload(ctx, m.get(), w->at("stem.b").data);
    m->stem_prelu = dev_upload(ctx, m.get(), w->at("stem.prelu").data);
    if (w->at("stem.prelu").data.size() != 64) return fr_fail(ctx, FR_ERR_MODEL, "stem.prelu must have 64 slopes");
    memcpy(m->stem_prelu_h, w->at("stem.prelu").data.data(), sizeof(m->stem_prelu_h));
  }
  const int layers[4][2] = {{3, 64}, {4, 128}, {14, 256}, {3, 512}};
  int cin = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < layers[l][0]; ++b) {
      const std::string p = "l" + std::to_string(l) + "." + std::to_string(b);
      BlockW bw;
      bw.cin = cin;
      bw.planes = layers[l][1];
      bw.stride = b == 0 ? 2 : 1;
      const fr_tensor& w1 = w->at(p + ".conv1.w");
      const std::vector<float>& sc = w->at(p + ".bn1.scale").data;
      const std::vector<float>& sh = w->at(p + ".bn1.shift").data;
      bw.w1 = dev_upload(ctx, m.get(), pack_conv3(w1, &sc));
      bw.b1 = dev_upload(ctx, m.get(), border_bias(w1, w->at(p + ".conv1.b").data, sh));
      bw.prelu = dev_upload(ctx, m.get(), w->at(p + ".prelu").data);
      bw.prelu_h = w->at(p + ".prelu").data;
      bw.w2 = dev_upload(ctx, m.get(), pack_conv3(w->at(p + ".conv2.w"), nullptr));
      std::vector<float> b2 = w->at(p + ".conv2.b").data;
      if (b == 0) {
        const fr_tensor& wd = w->at(p + ".ds.w");
        bw.wds = dev_upload(ctx, m.get(), to_bf16(wd.data));  // [planes][cin] already K-major
        const std::vector<float>& bd = w->at(p + ".ds.b").data;
        for (size_t i = 0; i < b2.size(); ++i) b2[i] += bd[i];
      }
      bw.b2 = dev_upload(ctx, m.get(), b2);
      if (!bw.w1 || !bw.b1 || !bw.prelu || !bw.w2 || !bw.b2)
        return fr_fail(ctx, FR_ERR_CUDA, "rec weight upload failed");
      m->blocks.push_back(bw);
      cin = bw.planes;
    }
  {
    // tail: out = fs * (Wfc * (s2 .* x + t2) + bfc) + ft, x in padded 8x8x512 cell layout
    const std::vector<float>& s2 = w->at("bn2.scale").data;
    const std::vector<float>& t2 = w->at("bn2.shift").data;
    const fr_tensor& fw = w->at("fc.w");
    const std::vector<float>& fb = w->at("fc.b").data;
    const std::vector<float>& fs = w->at("feat.scale").data;
    const std::vector<float>& ft = w->at("feat.shift").data;
    const int K = 64 * 512;
    std::vector<bf16> pw((size_t)512 * K, __float2bfloat16_rn(0.f));
    std::vector<float> pb(512);
    for (int o = 0; o < 512; ++o) {
      double acc = fb[o];
      for (int c = 0; c < 512; ++c)
        for (int hw = 0; hw < 49; ++hw) {
          const float wv = fw.data[(size_t)o * 25088 + (size_t)c * 49 + hw];
          acc += (double)wv * t2[c];
          const int h = hw / 7, x = hw % 7;
          pw[(size_t)o * K + (size_t)(h * 8 + x) * 512 + c] = __float2bfloat16_rn(fs[o] * wv * s2[c]);
        }
      pb[o] = (float)(fs[o] * acc + ft[o]);
    }
    m->fc_w = dev_upload(ctx, m.get(), pw);
    m->fc_b = dev_upload(ctx, m.get(), pb);
    if (!m->fc_w || !m->fc_b) return fr_fail(ctx, FR_ERR_CUDA, "fc weight upload failed");
  }
  if (cudaMalloc(&m->err_flag, sizeof(int)) != cudaSuccess)
    return fr_fail(ctx, FR_ERR_CUDA, "err flag alloc failed");
  cudaMemset(m->err_flag, 0, sizeof(int));
  ctx->rec = m.release();
  return FR_OK;
}

static void rec_free_plan(RecModel* m) {
  fr_alloc_epoch()++;          // cached CUDA graphs (capi.cu) point into the plan's buffers
  for (void* p : m->plan_allocs) cudaFree(p);
  m->plan_allocs.clear();
  m->bufs.clear();
  m->conv1.clear();
  m->conv2.clear();
  m->cap = 0;
}

void rec_model_destroy(fr_ctx* ctx) {
  RecModel* m = ctx->rec;
  if (!m) return;
  rec_free_plan(m);
  for (void* p : m->allocs) cudaFree(p);
  if (m->err_flag) cudaFree(m->err_flag);
  if (m->chw_stage) cudaFree(m->chw_stage);
  delete m;
  ctx->rec = nullptr;
}

namespace {

bool alloc_act(RecModel* m, Act& a, int cap, int H, int W, int C, cudaStream_t st) {
  a.H = H; a.W = W; a.C = C; a.Hp = H + 1; a.Wp = W + 1;
  void* p = nullptr;
  if (cudaMalloc(&p, a.bytes(cap)) != cudaSuccess) return false;
  cudaMemsetAsync(p, 0, a.bytes(cap), st);  // halo cells stay zero forever
  a.p = reinterpret_cast<bf16*>(p);
  m->plan_allocs.push_back(p);
  return true;
}

}  // namespace

// Build buffers, tensor maps and launch parameters for up to `cap` faces per call.
static int rec_build_plan(fr_ctx* ctx, int cap) {
  RecModel* m = ctx->rec;
  if (m->cap >= cap) return FR_OK;
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  rec_free_plan(m);
  cudaStream_t st = ctx->stream;
  bool ok = alloc_act(m, m->x0, cap, 112, 112, 64, st) && alloc_act(m, m->x0e, cap, 56, 56, 64, st);
  const size_t nb = m->blocks.size();
  m->bufs.resize(nb);
  m->conv1.resize(nb);
  m->conv2.resize(nb);
  int H = 112;  // spatial size of the current residual stream
  Act x = m->x0, xe = m->x0e;
  for (size_t i = 0; ok && i < nb; ++i) {
    const BlockW& bw = m->blocks[i];
    RecModel::BlockBufs& bb = m->bufs[i];
    const int Ho = H / bw.stride;
    const bool next_ds = (i + 1 < nb) && m->blocks[i + 1].stride == 2;
    if (bw.stride == 2)
      ok = ok && alloc_act(m, bb.h, cap, Ho, Ho, 4 * bw.planes, st);  // s2d cells
    else
      ok = ok && alloc_act(m, bb.h, cap, H, H, bw.planes, st);
    ok = ok && alloc_act(m, bb.out, cap, Ho, Ho, bw.planes, st);
    if (next_ds) ok = ok && alloc_act(m, bb.out_even, cap, Ho / 2, Ho / 2, bw.planes, st);
    if (!ok) break;
    // ---- conv1: 3x3 s1, x -> h (bias by border class, PReLU)
    ConvLaunch& c1 = m->conv1[i];
    memset(&c1.p, 0, sizeof(c1.p));
    c1.bn = pick_bn(bw.planes);
    fill_taps_s1(c1.p, x.Wp, bw.planes, bw.cin);
    c1.p.n_tiles_n = bw.planes / c1.bn;
    c1.p.H = H; c1.p.W = H; c1.p.Hp = x.Hp; c1.p.Wp = x.Wp;
    c1.p.cout = bw.planes;
    c1.p.bias = bw.b1; c1.p.bias_classes = 9;
    c1.p.prelu = bw.prelu;
    if (bw.prelu_h.size() <= 512 && env_flag("FR_TC_PRELU_PARAMS", 1)) {
      memcpy(c1.p.prelu_c, bw.prelu_h.data(), bw.prelu_h.size() * sizeof(float));
      c1.p.prelu_in_params = 1;
    }
    c1.p.out = bb.h.p;
    c1.p.err_flag = m->err_flag;
    c1.rows_per_img = x.Hp * x.Wp;
    if (bw.stride == 2) {
      c1.p.out_mode = tc::OUT_S2D;
      c1.p.Hp2 = bb.h.Hp; c1.p.Wp2 = bb.h.Wp;
    } else {
      c1.p.out_mode = tc::OUT_STD;
    }
    ok = ok && tc_make_map_2d(&c1.a0, x.p, x.rows(cap), x.C, x.C, tc::BM);
    c1.a1 = c1.a0;
    ok = ok && tc_setup_halo(c1, x.p, x.rows(cap), x.C, x.Wp);
    ok = ok && tc_make_map_2d(&c1.b0, bw.w1, (uint64_t)9 * bw.planes, bw.cin, bw.cin, c1.bn);
    c1.has_b_small = ok && c1.halo && c1.bn > 64 && tc_make_map_2d(&c1.b_small, bw.w1, (uint64_t)9 * bw.planes, bw.cin, bw.cin, 64);
    c1.b1 = c1.b0;
    c1.tma_store = ok && c1.halo && env_flag("FR_TC_TMASTORE", 1) && c1.p.out_mode == tc::OUT_STD && c1.p.n_tiles_n == 1 &&
                   c1.bn <= 128 && tc_make_map_2d(&c1.out_map, bb.h.p, bb.h.rows(cap), bb.h.C, bb.h.C, tc::BM);
    c1.two_cta = ok && c1.halo && c1.mt == 1 && two_cta_eligible(c1.bn, bw.cin) &&
                 tc_make_map_2d(&c1.b_half, bw.w1, (uint64_t)9 * bw.planes, bw.cin, bw.cin, c1.bn / 2) &&
                 tc_make_map_2d(&c1.b_half2, bw.w1, (uint64_t)9 * bw.planes, bw.cin, bw.cin, std::max(c1.bn / 4, 16)) &&
                 tc_make_map_2d(&c1.b_half4, bw.w1, (uint64_t)9 * bw.planes, bw.cin, bw.cin, std::max(c1.bn / 8, 16));
    // ---- conv2: 3x3 stride s (+ fused 1x1 shortcut conv) + residual -> out
    ConvLaunch& c2 = m->conv2[i];
    memset(&c2.p, 0, sizeof(c2.p));
    c2.bn = pick_bn(bw.planes);
    c2.p.n_tiles_n = bw.planes / c2.bn;
    c2.p.H = Ho; c2.p.W = Ho; c2.p.Hp = bb.out.Hp; c2.p.Wp = bb.out.Wp;
    c2.p.cout = bw.planes;
    c2.p.bias = bw.b2; c2.p.bias_classes = 1;
    c2.p.out = bb.out.p;
    c2.p.out_mode = tc::OUT_STD;
    c2.p.err_flag = m->err_flag;
    c2.rows_per_img = bb.out.Hp * bb.out.Wp;
    if (next_ds) {
      c2.p.out_even = bb.out_even.p;
      c2.p.Hp2 = bb.out_even.Hp; c2.p.Wp2 = bb.out_even.Wp;
    }
    if (bw.stride == 2) {
      fill_taps_s2(c2.p, bb.h.Wp, bw.planes, bw.planes);
      tc::Tap& t = c2.p.taps[9];
      t.a_src = 1; t.b_src = 1; t.a_row_shift = 0; t.a_col = 0; t.b_row = 0; t.nkb = bw.cin / 64;
      c2.p.num_taps = 10;
      ok = ok && tc_make_map_2d(&c2.a0, bb.h.p, bb.h.rows(cap), bb.h.C, bb.h.C, tc::BM);
      ok = ok && tc_make_map_2d(&c2.a1, xe.p, xe.rows(cap), xe.C, xe.C, tc::BM);
      ok = ok && tc_make_map_2d(&c2.b1, bw.wds, bw.planes, bw.cin, bw.cin, c2.bn);

    } else {
      fill_taps_s1(c2.p, bb.h.Wp, bw.planes, bw.planes);
      c2.p.residual = x.p;
      ok = ok && tc_make_map_2d(&c2.a0, bb.h.p, bb.h.rows(cap), bb.h.C, bb.h.C, tc::BM);
      c2.a1 = c2.a0;
      ok = ok && tc_setup_halo(c2, bb.h.p, bb.h.rows(cap), bb.h.C, bb.h.Wp);
    }
    ok = ok && tc_make_map_2d(&c2.b0, bw.w2, (uint64_t)9 * bw.planes, bw.planes, bw.planes, c2.bn);
    c2.has_b_small = ok && c2.halo && c2.bn > 64 && tc_make_map_2d(&c2.b_small, bw.w2, (uint64_t)9 * bw.planes, bw.planes, bw.planes, 64);
    c2.tma_store = ok && c2.halo && env_flag("FR_TC_TMASTORE", 1) && c2.p.out_mode == tc::OUT_STD && c2.p.n_tiles_n == 1 &&
                   c2.bn <= 128 && tc_make_map_2d(&c2.out_map, bb.out.p, bb.out.rows(cap), bb.out.C, bb.out.C, tc::BM);
    c2.two_cta = ok && c2.halo && c2.mt == 1 && two_cta_eligible(c2.bn, bw.planes) &&
                 tc_make_map_2d(&c2.b_half, bw.w2, (uint64_t)9 * bw.planes, bw.planes, bw.planes, c2.bn / 2) &&
                 tc_make_map_2d(&c2.b_half2, bw.w2, (uint64_t)9 * bw.planes, bw.planes, bw.planes, std::max(c2.bn / 4, 16)) &&
                 tc_make_map_2d(&c2.b_half4, bw.w2, (uint64_t)9 * bw.planes, bw.planes, bw.planes, std::max(c2.bn / 8, 16));
    if (bw.stride != 2) c2.b1 = c2.b0;
    x = bb.out;
    xe = bb.out_even;
    H = Ho;
  }
  if (ok) {
    void* p = nullptr;
    ok = cudaMalloc(&p, (size_t)cap * 512 * sizeof(float)) == cudaSuccess;
    if (ok) { m->fc_out = reinterpret_cast<float*>(p); m->plan_allocs.push_back(p); }
    void* q = nullptr;
    ok = ok && cudaMalloc(&q, (size_t)FC_SPLITS * cap * 512 * sizeof(float)) == cudaSuccess;
    if (ok) { m->fc_part = reinterpret_cast<float*>(q); m->plan_allocs.push_back(q); }
  }
  if (ok) {
    ConvLaunch& f = m->fc;
    memset(&f.p, 0, sizeof(f.p));
    f.bn = 256;
    f.p.num_taps = 1;
    f.p.taps[0].nkb = 64 * 512 / 64;
    f.p.n_tiles_n = 2;
    f.p.H = f.p.W = f.p.Hp = f.p.Wp = 1;
    f.p.cout = 512;
    f.p.bias = m->fc_b; f.p.bias_classes = 1;
    f.p.out_mode = tc::OUT_F32;
    f.p.out_f32 = m->fc_part;
    f.p.k_splits = FC_SPLITS;
    f.p.k_split_len = (64 * 512 / 64) / FC_SPLITS;
    f.p.split_stride = (long long)cap * 512;
    f.p.err_flag = m->err_flag;
    f.rows_per_img = 1;
    ok = ok && tc_make_map_2d(&f.a0, x.p, (uint64_t)cap, 64 * 512, 64 * 512, tc::BM);
    f.a1 = f.a0;
    ok = ok && tc_make_map_2d(&f.b0, m->fc_w, 512, 64 * 512, 64 * 512, 256);
    f.b1 = f.b0;
  }
  if (!ok) {
    rec_free_plan(m);
    return fr_fail(ctx, FR_ERR_CUDA, "rec plan allocation / tensor map encode failed");
  }
  m->cap = cap;
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return FR_OK;
}

// The residual stream of layer 0 is 1.6 MB per face at 112^2 (0.4 MB at 56^2): at the bench's 512
// faces those tensors are 213-835 MB, stream through HBM between launches, and the launches that
// touch them are DRAM-bound (the stride-2 conv of block 0 reads 1.07 GB).  Running the stem and the
// first FR_REC_CHUNK_BLOCKS blocks depth-first over chunks of FR_REC_CHUNK faces keeps each chunk's
// tensors (<= 100 MB) resident in the 126 MB L2: every chunk reuses the same rows [0, chunk) of the
// plan's buffers, and only the last chunked block writes its output at the chunk's offset in the
// full-batch tensor.  29 faces: 29 * 57 * 57 / 128 = 736.1 -> 737 tiles = 4.98 waves of 148 CTAs.
// Same kernels, same per-element accumulation order: results are bit-identical to the unchunked run.
static int rec_chunk_faces() { static const int v = env_flag("FR_REC_CHUNK", 0); return v; }
static int rec_chunk_blocks() { static const int v = env_flag("FR_REC_CHUNK_BLOCKS", 3); return v; }

static int rec_launch_stem(fr_ctx* ctx, const uint8_t* d_crops, int n) {
  RecModel* m = ctx->rec;
  ctx->stage_begin(FR_STAGE_STEM);
  static const bool simt_stem = getenv("FR_STEM_SIMT") != nullptr;   // A/B switch: the CUDA-core stem
  if (simt_stem) {
    const long long threads = (long long)n * REC * (REC / 4);
    stem_kernel<true><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(
        d_crops, n, m->stem_w, m->stem_b, m->stem_prelu, m->x0.p, m->x0e.p);
  } else {
    StemSlopes sl;
    memcpy(sl.v, m->stem_prelu_h, sizeof(sl.v));
    stem_mma_kernel<<<(unsigned)(n * (REC / (STEM_WARPS * STEM_GROUPS))), STEM_WARPS * 32, 0, ctx->stream>>>(
        d_crops, m->stem_bfrag, sl, m->x0.p, m->x0e.p);
  }
  ctx->stage_end();
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

// stem + blocks [0, nb) for faces [f0, f0 + nf), computed in rows [0, nf) of the plan's buffers;
// block nb-1 writes its output (and the even-pixel copy) at face offset f0
static int rec_run_front_chunk(fr_ctx* ctx, const uint8_t* d_crops, int f0, int nf, int nb) {
  RecModel* m = ctx->rec;
  FR_CHECK(rec_launch_stem(ctx, d_crops + (size_t)f0 * REC * REC * 3, nf));
  ctx->stage_begin(FR_STAGE_TRUNK);
  for (int i = 0; i < nb; ++i) {
    FR_CHECK(tc_launch(ctx, m->conv1[i], nf * m->conv1[i].rows_per_img));
    ConvLaunch& c2 = m->conv2[i];
    bf16* const out0 = c2.p.out;
    bf16* const even0 = c2.p.out_even;
    if (i == nb - 1) {
      const RecModel::BlockBufs& bb = m->bufs[i];
      c2.p.out = out0 + bb.out.rows(f0) * bb.out.C;
      if (even0) c2.p.out_even = even0 + bb.out_even.rows(f0) * bb.out_even.C;
    }
    const int s = tc_launch(ctx, c2, nf * c2.rows_per_img);
    c2.p.out = out0;
    c2.p.out_even = even0;
    FR_CHECK(s);
  }
  ctx->stage_end();
  return FR_OK;
}

static int rec_run_trunk(fr_ctx* ctx, int n, float* d_out_raw, size_t first_block = 0) {
  RecModel* m = ctx->rec;
  for (size_t i = first_block; i < m->blocks.size(); ++i) {
    FR_CHECK(tc_launch(ctx, m->conv1[i], n * m->conv1[i].rows_per_img));
    FR_CHECK(tc_launch(ctx, m->conv2[i], n * m->conv2[i].rows_per_img));
  }
  FR_CHECK(tc_launch(ctx, m->fc, n));
  float* raw = d_out_raw ? d_out_raw : m->fc_out;
  const size_t elems = (size_t)n * 512;
  fc_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, ctx->stream>>>(m->fc_part, elems, (size_t)m->cap * 512, raw);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

static int rec_plan_cap(int n) {
  int cap = 128;
  while (cap < n) cap *= 2;
  return cap;
}

int rec_forward_crops(fr_ctx* ctx, const uint8_t* d_crops, int n, float* d_out_raw,
                      float* d_out_norm, const int* d_valid) {
  RecModel* m = ctx->rec;
  if (!m) return fr_fail(ctx, FR_ERR_NOT_LOADED, "Model not loaded!");
  if (n <= 0) return FR_OK;
  FR_CHECK(rec_build_plan(ctx, rec_plan_cap(n)));
  float* raw = d_out_raw ? d_out_raw : m->fc_out;
  const int chunk = rec_chunk_faces();
  const int cb = std::min<int>(rec_chunk_blocks(), (int)m->blocks.size());
  size_t first_block = 0;
  if (chunk > 0 && cb > 0 && n > chunk) {
    for (int f0 = 0; f0 < n; f0 += chunk) FR_CHECK(rec_run_front_chunk(ctx, d_crops, f0, std::min(chunk, n - f0), cb));
    first_block = (size_t)cb;
  } else {
    FR_CHECK(rec_launch_stem(ctx, d_crops, n));
  }
  ctx->stage_begin(FR_STAGE_TRUNK);
  FR_CHECK(rec_run_trunk(ctx, n, raw, first_block));
  ctx->stage_end();
  if (d_out_norm) {
    ctx->stage_begin(FR_STAGE_L2NORM);
    FR_CHECK(k_l2_normalize(ctx, raw, n, 512, d_out_norm, d_valid));
    ctx->stage_end();
  }
  return FR_OK;
}

int rec_forward_chw(fr_ctx* ctx, const float* d_chw, int n, float* d_out_raw) {
  RecModel* m = ctx->rec;
  if (!m) return fr_fail(ctx, FR_ERR_NOT_LOADED, "Model not loaded!");
  if (n <= 0) return FR_OK;
  FR_CHECK(rec_build_plan(ctx, rec_plan_cap(n)));
  const long long threads = (long long)n * REC * (REC / 4);
  stem_kernel<false><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(
      d_chw, n, m->stem_w, m->stem_b, m->stem_prelu, m->x0.p, m->x0e.p);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return rec_run_trunk(ctx, n, d_out_raw);
}

// Copy an activation back as fp32 NCHW.  tap 0 = stem; 2*i-1 = h of block i; 2*i = out of block i.
int rec_tap(fr_ctx* ctx, int tap, int n, float* h_out, size_t out_elems) {
  RecModel* m = ctx->rec;
  if (!m || m->cap < n) return fr_fail(ctx, FR_ERR_INVALID_ARG, "rec_tap: no forward has run");
  Act a;
  bool s2d = false;
  if (tap == 0) {
    a = m->x0;
  } else {
    const int blk = (tap - 1) / 2;
    if (blk < 0 || blk >= (int)m->blocks.size()) return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad tap");
    if (tap & 1) { a = m->bufs[blk].h; s2d = m->blocks[blk].stride == 2; }
    else a = m->bufs[blk].out;
  }
  const int C = s2d ? a.C / 4 : a.C;
  const int H = s2d ? a.H * 2 : a.H, W = s2d ? a.W * 2 : a.W;
  if (out_elems != (size_t)n * C * H * W) return fr_fail(ctx, FR_ERR_CAPACITY, "rec_tap: size mismatch");
  std::vector<bf16> host(a.rows(n) * a.C);
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  FR_CUDA_OK(ctx, cudaMemcpy(host.data(), a.p, host.size() * 2, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          size_t src;
          if (s2d)
            src = (((size_t)(i * a.Hp + (y >> 1)) * a.Wp + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * C + c;
          else
            src = ((size_t)(i * a.Hp + y) * a.Wp + x) * C + c;
          h_out[(((size_t)i * C + c) * H + y) * W + x] = __bfloat162float(host[src]);
        }
  return FR_OK;
}

// Unit-test hook: one 3x3 stride-1 (or 1x1) convolution through the tcgen05 kernel.
int rec_test_conv(fr_ctx* ctx, const float* x, int n, int cin, int h, int w, const float* wgt,
                  int cout, int ksize, int stride, const float* pre_scale,
                  const float* pre_shift, const float* bias, const float* prelu,
                  const float* residual, float* y) {
  if (stride != 1 || (ksize != 3 && ksize != 1) || cin % 64 || cout % 64 ||
      (cout > 256 && cout % 256))
    return fr_fail(ctx, FR_ERR_UNSUPPORTED, "fr_test_conv: unsupported shape");
  if (!get_encode_fn()) return fr_fail(ctx, FR_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const int Hp = h + 1, Wp = w + 1;
  const size_t rows = (size_t)n * Hp * Wp;
  std::vector<bf16> xin(rows * cin, __float2bfloat16_rn(0.f));
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < cin; ++c)
      for (int yy = 0; yy < h; ++yy)
        for (int xx = 0; xx < w; ++xx)
          xin[((size_t)(i * Hp + yy) * Wp + xx) * cin + c] =
              __float2bfloat16_rn(x[(((size_t)i * cin + c) * h + yy) * w + xx]);
  fr_tensor wt;
  wt.dims = {cout, cin, ksize, ksize};
  wt.data.assign(wgt, wgt + (size_t)cout * cin * ksize * ksize);
  std::vector<float> sc, sh(cin, 0.f), bs(cout, 0.f);
  if (pre_scale) sc.assign(pre_scale, pre_scale + cin);
  if (pre_shift) sh.assign(pre_shift, pre_shift + cin);
  if (bias) bs.assign(bias, bias + cout);
  std::vector<bf16> pw = pack_conv3(wt, pre_scale ? &sc : nullptr);
  std::vector<float> bt;
  if (ksize == 3) {
    bt = border_bias(wt, bs, sh);
  } else {
    bt = bs;
    for (int a = 0; a < cout; ++a) {
      double s = 0;
      for (int c = 0; c < cin; ++c) s += (double)wgt[(size_t)a * cin + c] * sh[c];
      bt[a] += (float)s;
    }
  }
  std::vector<bf16> res;
  if (residual) {
    res.assign(rows * cout, __float2bfloat16_rn(0.f));
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < cout; ++c)
        for (int yy = 0; yy < h; ++yy)
          for (int xx = 0; xx < w; ++xx)
            res[((size_t)(i * Hp + yy) * Wp + xx) * cout + c] =
                __float2bfloat16_rn(residual[(((size_t)i * cout + c) * h + yy) * w + xx]);
  }
  bf16 *d_x = nullptr, *d_w = nullptr, *d_res = nullptr, *d_y = nullptr;
  float *d_b = nullptr, *d_p = nullptr;
  int* d_err = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_x); cudaFree(d_w); cudaFree(d_res); cudaFree(d_y); cudaFree(d_b); cudaFree(d_p);
    cudaFree(d_err);
  };
  bool ok = cudaMalloc(&d_x, xin.size() * 2) == cudaSuccess &&
            cudaMalloc(&d_w, pw.size() * 2) == cudaSuccess &&
            cudaMalloc(&d_y, rows * cout * 2) == cudaSuccess &&
            cudaMalloc(&d_b, bt.size() * 4) == cudaSuccess &&
            cudaMalloc(&d_err, 4) == cudaSuccess;
  if (ok && prelu) ok = cudaMalloc(&d_p, cout * 4) == cudaSuccess;
  if (ok && residual) ok = cudaMalloc(&d_res, res.size() * 2) == cudaSuccess;
  if (!ok) { cleanup(); return fr_fail(ctx, FR_ERR_CUDA, "fr_test_conv: alloc failed"); }
  cudaMemcpy(d_x, xin.data(), xin.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_w, pw.data(), pw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_y, 0, rows * cout * 2);
  cudaMemset(d_err, 0, 4);
  if (prelu) cudaMemcpy(d_p, prelu, cout * 4, cudaMemcpyHostToDevice);
  if (residual) cudaMemcpy(d_res, res.data(), res.size() * 2, cudaMemcpyHostToDevice);
  ConvLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  L.bn = pick_bn(cout);
  if (ksize == 3) {
    fill_taps_s1(L.p, Wp, cout, cin);
    L.p.bias_classes = 9;
  } else {
    L.p.num_taps = 1;
    L.p.taps[0].nkb = cin / 64;
    L.p.bias_classes = 1;
  }
  L.p.n_tiles_n = cout / L.bn;
  L.p.H = h; L.p.W = w; L.p.Hp = Hp; L.p.Wp = Wp;
  L.p.cout = cout;
  L.p.bias = d_b;
  L.p.prelu = d_p;
  L.p.residual = d_res;
  L.p.out = d_y;
  L.p.out_mode = tc::OUT_STD;
  L.p.err_flag = d_err;
  ok = tc_make_map_2d(&L.a0, d_x, rows, cin, cin, tc::BM) &&
       tc_make_map_2d(&L.b0, d_w, (uint64_t)ksize * ksize * cout, cin, cin, L.bn);
  L.a1 = L.a0;
  L.b1 = L.b0;
  if (ok && ksize == 3) ok = tc_setup_halo(L, d_x, rows, cin, Wp);
  L.has_b_small = ok && L.halo && L.bn > 64 && tc_make_map_2d(&L.b_small, d_w, (uint64_t)ksize * ksize * cout, cin, cin, 64);
  L.tma_store = ok && L.halo && env_flag("FR_TC_TMASTORE", 1) && L.p.n_tiles_n == 1 && L.bn <= 128 &&
                tc_make_map_2d(&L.out_map, d_y, rows, cout, cout, tc::BM);
  L.two_cta = ok && L.halo && L.mt == 1 && two_cta_eligible(L.bn, cin) &&
              tc_make_map_2d(&L.b_half, d_w, (uint64_t)9 * cout, cin, cin, L.bn / 2) &&
              tc_make_map_2d(&L.b_half2, d_w, (uint64_t)9 * cout, cin, cin, std::max(L.bn / 4, 16)) &&
              tc_make_map_2d(&L.b_half4, d_w, (uint64_t)9 * cout, cin, cin, std::max(L.bn / 8, 16));
  int status = FR_OK;
  if (!ok) status = fr_fail(ctx, FR_ERR_CUDA, "fr_test_conv: tensor map encode failed");
  if (status == FR_OK) status = tc_launch(ctx, L, (int)rows);
  if (status == FR_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) status = fr_fail(ctx, FR_ERR_CUDA, std::string("fr_test_conv: ") + cudaGetErrorString(e));
  }
  if (status == FR_OK) {
    std::vector<bf16> yo(rows * cout);
    cudaMemcpy(yo.data(), d_y, yo.size() * 2, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < cout; ++c)
        for (int yy = 0; yy < h; ++yy)
          for (int xx = 0; xx < w; ++xx)
            y[(((size_t)i * cout + c) * h + yy) * w + xx] =
                __bfloat162float(yo[((size_t)(i * Hp + yy) * Wp + xx) * cout + c]);
  }
  cleanup();
  return status;
}

int k_l2_normalize(fr_ctx* ctx, const float* d_in, int n, int dim, float* d_out,
                   const int* d_valid) {
  if (n <= 0) return FR_OK;
  l2_normalize_kernel<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(d_in, n, dim, d_out, d_valid);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

int k_compare_batch(fr_ctx* ctx, const float* d_a, const float* d_b, int n, int dim, float* d_out) {
  if (n <= 0) return FR_OK;
  compare_batch_kernel<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(d_a, d_b, n, dim, d_out);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}
