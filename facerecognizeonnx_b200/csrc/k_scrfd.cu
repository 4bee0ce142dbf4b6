// K2: SCRFD det_500m forward -- the replacement for `session_->Run` in FaceDetector::detect
// (reference src/face_detector.cpp:170-183).  Architecture per InsightFace scrfd_500m_bnkps
// (depthwise-separable backbone 16/40/72/152/288, PAFPN 16 ch, per-stride heads 64 ch,
// 2 anchors), BN folded into conv+bias, sigmoid on the score heads.
//
// Boxes must hold 1e-3 px against an fp32 CPU engine, which rules out bf16 (2^-9) and plain
// tf32 (2^-11) products.  Every pointwise / dense convolution therefore runs on the 5th-gen
// tensor cores as a THREE-TERM TF32 SPLIT: x = x_hi + x_lo with x_hi = rna_tf32(x), and
//     x * w  ~=  x_lo*w_hi + x_hi*w_lo + x_hi*w_hi          (dropped term <= 2^-22 |x w|)
// accumulated in fp32 in TMEM -- fp32-grade results at tensor-core rate.
//
// One kernel (`sep_gemm_kernel`) serves all 33 GEMM-shaped layers.  Activations are fp32 NHWC
// (K = channels contiguous).  Per 128-pixel tile and per 32-channel K block:
//   loader warps (8, two groups working on alternating pipeline stages): gather the A operand
//       straight from global memory -- plain (1x1), im2col (dense 3x3) or with the depthwise
//       3x3 + bias + ReLU evaluated on the fly (the depthwise result never exists in HBM) --
//       split it into tf32 hi/lo and store both as K-major SWIZZLE_128B tiles in shared memory,
//       then fence.proxy.async + mbarrier arrive;
//   warp 0 : TMA of the (host-pre-split) hi/lo weight tiles into the same stage;
//   warp 1 : one lane issues 3 x tcgen05.mma.kind::tf32 (M=128, N=16..160, K=8) per K step;
//   warps 2-5: epilogue, tcgen05.ld -> bias / ReLU / top-down upsample-add / accumulate ->
//       fp32 NHWC float4 stores, or the anchor-major score (sigmoid) / bbox / kps scatter.
// TMEM accumulators are double buffered, so gather, MMA and epilogue of consecutive tiles overlap.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "tc_gemm.cuh"

bool tc_make_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t pitch_elems, uint32_t box_rows);   // k_iresnet.cu

namespace {

using tc::make_smem_desc;
using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tma_load_2d;

constexpr int DET = FR_DET_SIZE;
constexpr int TM = 128;                 // pixels per tile (MMA M)
constexpr int A_BYTES = TM * 128;       // one 128 x 32 fp32 operand tile
constexpr int MAX_STAGES = 6;
constexpr int LD_GROUPS = 2;            // loader groups, working on alternating stages
constexpr int LD_WARPS = 8;
constexpr int TG = LD_WARPS * 32 / LD_GROUPS;   // threads per loader group
constexpr int ITEMS = TM * 8 / TG;              // (row, 16-byte chunk) items per thread per K block
constexpr int ROW_STEP = TG / 8;
constexpr int FIRST_LD_WARP = 6;
constexpr int THREADS = (FIRST_LD_WARP + LD_WARPS) * 32;
static_assert(ROW_STEP % 8 == 0, "row & 7 must be constant per thread");

enum { LD_PW = 0, LD_DW = 1, LD_IM2COL = 2 };
enum { EPI_STD = 0, EPI_HEAD = 1 };

struct SepParams {
  const float* in;      // fp32 NHWC [n][hin][win][cin]
  float* out;           // fp32 NHWC [n][hout][wout][cout]
  const float* dw_w;    // [9][cin] depthwise weights, tap major
  const float* dw_b;    // [cin]
  const float* bias;    // [npad_total]
  const float* add_up;  // optional [n][hout/2][wout/2][cout], nearest-upsampled and added
  float* score;
  float* bbox;
  float* kps;
  int cin, cout, kdim;
  int hin, win, hout, wout, stride;
  int mode, epi, relu, accumulate;
  int nkb;              // 32-wide K blocks
  int nt;               // MMA N (per N tile)
  int n_tiles_n;
  int npad_total;       // rows of the hi half of the packed weights
  int total_px;         // n * hout * wout
  int tiles_x, tiles_y; // 8 x 16 spatial tiles per image (0 = flattened pixel order)
  int num_m_tiles;
  int stages;
  int tmem_cols;
  int nbig;             // hi*hi accumulators per tile (k steps alternate between them)
  int acc_stages;       // 2 = TMEM double buffered, 1 = single
  int* err_flag;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, uint32_t off, const float4& v) {
  float4 h, l;
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  // the residual is rounded (not left to the tensor core's truncation) to tf32 as well
  l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

__device__ __forceinline__ void fma4(float4& acc, const float4& v, const float4& w) {
  acc.x = fmaf(v.x, w.x, acc.x);
  acc.y = fmaf(v.y, w.y, acc.y);
  acc.z = fmaf(v.z, w.z, acc.z);
  acc.w = fmaf(v.w, w.w, acc.w);
}

// kind::tf32 instruction descriptor: tf32 x tf32 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 accumulator columns of one row: hi*hi partial sum(s) + the cross-term accumulator, added
// in fp32 with round-to-nearest (small + (big0 + big1)).
__device__ __forceinline__ void ld_sum16(uint32_t taddr, uint32_t small_off, int nbig, uint32_t nt, uint32_t (&v)[16]) {
  uint32_t s[16];
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + small_off, s);
  if (nbig == 2) {
    uint32_t b1[16];
    tmem_ld16(taddr + nt, b1);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(b1[i]));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(s[i]));
}

// tile row -> output pixel.  2-D mode: 8 x 16 pixel patches (3x3 stencils of a patch overlap in
// L1 instead of re-reading whole image rows); flattened mode for the 40^2 / 20^2 maps.
__device__ __forceinline__ bool tile_pixel(const SepParams& p, int m_tile, int row, int& n, int& oy, int& ox) {
  if (p.tiles_x > 0) {
    const int per_img = p.tiles_x * p.tiles_y;
    n = m_tile / per_img;
    const int t = m_tile - n * per_img;
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    oy = ty * 8 + (row >> 4);
    ox = tx * 16 + (row & 15);
    return true;
  }
  const int m = m_tile * TM + row;
  if (m >= p.total_px) {
    n = oy = ox = 0;
    return false;
  }
  const int hw = p.hout * p.wout;
  n = m / hw;
  const int r = m - n * hw;
  oy = r / p.wout;
  ox = r - oy * p.wout;
  return true;
}

__global__ void __launch_bounds__(THREADS, 1)
sep_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ SepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = p.nt * 128;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tfull = empty + MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], TG + 1);      // loader group threads + the weight TMA's expect_tx arrive
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::prefetch_tmap(&tmW);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = p.num_m_tiles * p.n_tiles_n;

  if (warp == 0) {
    // ------------------------------------------------------------- weight TMA
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles_n) * p.nt;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int stage = it % p.stages;
          const uint32_t phase = (it / p.stages) & 1u;
          mbar_wait(&empty[stage], phase ^ 1u, p.err_flag);
          mbar_expect_tx(&full[stage], 2u * (uint32_t)b_bytes);
          uint8_t* sb = smem + stage * stage_bytes + 2 * A_BYTES;
          tma_load_2d(sb, &tmW, &full[stage], kb * 32, n0);
          tma_load_2d(sb + b_bytes, &tmW, &full[stage], kb * 32, p.npad_total + n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TM, p.nt);
      uint32_t it = 0, tt = 0;
      const uint32_t slot_cols = (uint32_t)((p.nbig + 1) * p.nt);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tt) {
        const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
        const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        // The tensor core truncates when it adds a K=8 product group into the fp32 accumulator,
        // a bias of ~2^-24 |D| per MMA.  So: the two small cross terms go to their own
        // accumulator (their truncation is 2^-11 smaller), and the hi*hi products alternate
        // between nbig accumulators; the epilogue adds the partial sums with round-to-nearest.
        const uint32_t d_big = tmem_base + acc * slot_cols;
        const uint32_t d_small = d_big + (uint32_t)(p.nbig * p.nt);
        uint32_t ks_total = 0;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int stage = it % p.stages;
          const uint32_t phase = (it / p.stages) & 1u;
          mbar_wait(&full[stage], phase, p.err_flag);
          tc_fence_after();
          const uint8_t* sa = smem + stage * stage_bytes;
          const uint64_t ahi = make_smem_desc(sa);
          const uint64_t alo = make_smem_desc(sa + A_BYTES);
          const uint64_t bhi = make_smem_desc(sa + 2 * A_BYTES);
          const uint64_t blo = make_smem_desc(sa + 2 * A_BYTES + b_bytes);
          const int ksteps = min(4, (p.kdim - kb * 32 + 7) >> 3);   // K = 8 tf32 per MMA
          for (int k = 0; k < ksteps; ++k, ++ks_total) {
            const uint64_t o = (uint64_t)(k * 2);   // +32 bytes inside the 128-byte swizzle row
            mma_tf32(d_small, alo + o, bhi + o, idesc, ks_total ? 1u : 0u);
            mma_tf32(d_small, ahi + o, blo + o, idesc, 1u);
            const uint32_t which = p.nbig == 2 ? (ks_total & 1u) : 0u;
            mma_tf32(d_big + which * (uint32_t)p.nt, ahi + o, bhi + o, idesc, ks_total >= (uint32_t)p.nbig ? 1u : 0u);
          }
          tc_commit(&empty[stage]);
        }
        tc_commit(&tfull[acc]);
      }
    }
  } else if (warp < FIRST_LD_WARP) {
    // ------------------------------------------------------------- epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint32_t tt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tt) {
      const int m_tile = tile / p.n_tiles_n;
      const int n0 = (tile % p.n_tiles_n) * p.nt;
      int n, oy, ox;
      const bool valid = tile_pixel(p, m_tile, row, n, oy, ox);
      const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (uint32_t)((p.nbig + 1) * p.nt) + ((uint32_t)(q * 32) << 16);
      // kdim <= 8 * nbig would leave the second hi*hi accumulator unwritten; not the case here
      const uint32_t small_off = (uint32_t)(p.nbig * p.nt);
      if (p.epi == EPI_HEAD) {
        // 30 channels of one pixel: [0,2) score (sigmoid), [2,10) bbox, [10,30) kps; the export's
        // anchor-major layout is anchor = pixel*2 + a, i.e. contiguous per pixel
        uint32_t v0[16], v1[16];
        ld_sum16(taddr, small_off, p.nbig, (uint32_t)p.nt, v0);
        ld_sum16(taddr + 16, small_off, p.nbig, (uint32_t)p.nt, v1);
        if (valid) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = __uint_as_float(v0[i]) + __ldg(p.bias + i);
            f[16 + i] = __uint_as_float(v1[i]) + __ldg(p.bias + 16 + i);
          }
          const size_t hw = (size_t)p.hout * p.wout;
          const size_t pix = (size_t)oy * p.wout + ox;
          float2 sc;
          sc.x = 1.0f / (1.0f + expf(-f[0]));
          sc.y = 1.0f / (1.0f + expf(-f[1]));
          *reinterpret_cast<float2*>(p.score + ((size_t)n * hw + pix) * 2) = sc;
          float4* bb = reinterpret_cast<float4*>(p.bbox + ((size_t)n * hw + pix) * 8);
          bb[0] = make_float4(f[2], f[3], f[4], f[5]);
          bb[1] = make_float4(f[6], f[7], f[8], f[9]);
          float4* kp = reinterpret_cast<float4*>(p.kps + ((size_t)n * hw + pix) * 20);
#pragma unroll
          for (int i = 0; i < 5; ++i) kp[i] = make_float4(f[10 + 4 * i], f[11 + 4 * i], f[12 + 4 * i], f[13 + 4 * i]);
        }
      } else {
        const size_t pix_off = valid ? ((size_t)(n * p.hout + oy) * p.wout + ox) * p.cout : 0;
        const float* up = nullptr;
        if (p.add_up && valid)
          up = p.add_up + ((size_t)(n * (p.hout >> 1) + (oy >> 1)) * (p.wout >> 1) + (ox >> 1)) * p.cout;
#pragma unroll 1
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
          uint32_t v[16];
          ld_sum16(taddr + c0, small_off, p.nbig, (uint32_t)p.nt, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = n0 + c0 + 4 * j;
              if (c < p.cout) {
                const float4 b4 = ldg4(p.bias + c);
                float4 f = make_float4(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y,
                                       __uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
                if (p.relu) {
                  f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
                }
                if (up) {
                  const float4 u = ldg4(up + c);
                  f.x += u.x; f.y += u.y; f.z += u.z; f.w += u.w;
                }
                float4* o = reinterpret_cast<float4*>(p.out + pix_off + c);
                if (p.accumulate) {
                  const float4 e = *o;
                  f.x += e.x; f.y += e.y; f.z += e.z; f.w += e.w;
                }
                *o = f;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  } else {
    // ------------------------------------------------------------- A-operand gather
    const int lw = warp - FIRST_LD_WARP;
    const int grp = lw % LD_GROUPS;
    const int lt = (lw / LD_GROUPS) * 32 + lane;     // 0 .. TG-1 within the group
    const int chunk = lt & 7;                        // 16-byte chunk = 4 channels of the K block
    const int r0 = lt >> 3;
    const uint32_t off0 = (uint32_t)r0 * 128u + (uint32_t)((chunk ^ (r0 & 7)) << 4);   // SWIZZLE_128B
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles_n;
      bool have = false;
      int pix[ITEMS];
      for (int kb = 0; kb < p.nkb; ++kb, ++it) {
        if ((int)(it % LD_GROUPS) != grp) continue;
        if (!have) {
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            int n, oy, ox;
            const bool ok = tile_pixel(p, m_tile, r0 + j * ROW_STEP, n, oy, ox);
            pix[j] = ok ? ((n << 20) | (oy << 10) | ox) : -1;
          }
          have = true;
        }
        const int stage = it % p.stages;
        const uint32_t phase = (it / p.stages) & 1u;
        mbar_wait(&empty[stage], phase ^ 1u, p.err_flag);
        uint8_t* hi = smem + stage * stage_bytes;
        uint8_t* lo = hi + A_BYTES;
        const int k = kb * 32 + chunk * 4;
        if (p.mode == LD_DW) {
          const bool kv = k < p.cin;
          float4 w[9], b4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = kv ? ldg4(p.dw_w + t * p.cin + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (kv) b4 = ldg4(p.dw_b + k);
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kv && pix[j] >= 0) {
              const int n = pix[j] >> 20, oy = (pix[j] >> 10) & 1023, ox = pix[j] & 1023;
              const float* base = p.in + (size_t)n * p.hin * p.win * p.cin + k;
              const int iy0 = oy * p.stride - 1, ix0 = ox * p.stride - 1;
              float4 v[9];
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                  const int iy = iy0 + r, ix = ix0 + s;
                  const bool ok = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win;
                  v[r * 3 + s] = ok ? ldg4(base + ((size_t)iy * p.win + ix) * p.cin) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              acc = b4;
#pragma unroll
              for (int t = 0; t < 9; ++t) fma4(acc, v[t], w[t]);
              acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
              acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
            }
            split_store(hi, lo, off0 + (uint32_t)(j * ROW_STEP) * 128u, acc);
          }
        } else if (p.mode == LD_PW) {
          const bool kv = k < p.cin;
          float4 v[ITEMS];
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kv && pix[j] >= 0) {
              const int n = pix[j] >> 20, oy = (pix[j] >> 10) & 1023, ox = pix[j] & 1023;
              v[j] = ldg4(p.in + ((size_t)(n * p.hin + oy) * p.win + ox) * p.cin + k);
            }
          }
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) split_store(hi, lo, off0 + (uint32_t)(j * ROW_STEP) * 128u, v[j]);
        } else {
          // dense 3x3, K index = tap * cin + c (a 16-byte chunk never straddles a tap: cin % 4 == 0)
          const int tap = k / p.cin;
          const int c = k - tap * p.cin;
          const bool kv = tap < 9;
          const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          float4 v[ITEMS];
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kv && pix[j] >= 0) {
              const int n = pix[j] >> 20, oy = (pix[j] >> 10) & 1023, ox = pix[j] & 1023;
              const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
              if (iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win)
                v[j] = ldg4(p.in + ((size_t)(n * p.hin + iy) * p.win + ix) * p.cin + c);
            }
          }
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) split_store(hi, lo, off0 + (uint32_t)(j * ROW_STEP) * 128u, v[j]);
        }
        // generic-proxy stores -> visible to the tensor core (async proxy), then publish
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full[stage]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Stem: dense 3x3 stride-2 conv 3 -> 16 + bias + ReLU on the bf16 planar input produced by K1
// (K = 27: CUDA cores; the layer is bound by its 6.5 MB / frame fp32 NHWC write).
// One thread = one output pixel x 16 channels = one 64-byte NHWC record.
__global__ void __launch_bounds__(256)
stem_conv_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                 const float* __restrict__ b, int n_img) {
  __shared__ __align__(16) float sw[27 * 16];
  __shared__ __align__(16) float sb[16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 16) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  constexpr int HO = DET / 2;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)n_img * HO * HO) return;
  const int n = (int)(gid / (HO * HO));
  const int pp = (int)(gid - (size_t)n * HO * HO);
  const int oy = pp / HO, ox = pp - oy * HO;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = sb[i];
  const __nv_bfloat16* ip = in + (size_t)n * 3 * DET * DET;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 - 1 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * 2 - 1 + s;
        float v = 0.f;
        if (iy >= 0 && iy < DET && ix >= 0 && ix < DET) v = __bfloat162float(ip[((size_t)c * DET + iy) * DET + ix]);
        const float4* wp = reinterpret_cast<const float4*>(&sw[((c * 3 + r) * 3 + s) * 16]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w4 = wp[i];
          acc[4 * i] = fmaf(v, w4.x, acc[4 * i]);
          acc[4 * i + 1] = fmaf(v, w4.y, acc[4 * i + 1]);
          acc[4 * i + 2] = fmaf(v, w4.z, acc[4 * i + 2]);
          acc[4 * i + 3] = fmaf(v, w4.w, acc[4 * i + 3]);
        }
      }
    }
  float4* op = reinterpret_cast<float4*>(out + gid * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    op[i] = make_float4(fmaxf(acc[4 * i], 0.f), fmaxf(acc[4 * i + 1], 0.f), fmaxf(acc[4 * i + 2], 0.f),
                        fmaxf(acc[4 * i + 3], 0.f));
}

}  // namespace

struct PackedConv {        // device-side packed parameters of one (fused) layer
  float* wpack = nullptr;  // [2][npad_total][kpad]: tf32 hi half, then the residual lo half
  float* bias = nullptr;   // [npad_total]
  float* dw_w = nullptr;   // [9][cin]  (dw-separable only)
  float* dw_b = nullptr;   // [cin]
  CUtensorMap tmW;
  int cin = 0, cout = 0, kdim = 0, nkb = 0, nt = 0, n_tiles_n = 1, npad_total = 0, mode = LD_PW;
};

struct DetModel {
  std::map<std::string, PackedConv> conv;
  float* stem_w = nullptr;
  float* stem_b = nullptr;
  std::vector<void*> allocs;
  int cap = 0;
  std::vector<void*> act_allocs;
  // activations (fp32 NHWC)
  float *a_stem = nullptr, *a_b0 = nullptr;
  std::vector<float*> a_stage;       // per dw-separable block output
  float* lat[3] = {nullptr, nullptr, nullptr};
  float* inter[3] = {nullptr, nullptr, nullptr};
  float* pout[3] = {nullptr, nullptr, nullptr};
  float* tw0[3] = {nullptr, nullptr, nullptr};
  float* tw1[3] = {nullptr, nullptr, nullptr};
  float* score[3] = {nullptr, nullptr, nullptr};
  float* bbox[3] = {nullptr, nullptr, nullptr};
  float* kps[3] = {nullptr, nullptr, nullptr};
  int* err_flag = nullptr;
  int num_sms = 148;
};

namespace {

const int kStages[4][2] = {{2, 40}, {3, 72}, {2, 152}, {6, 288}};

float* upload(DetModel* m, const std::vector<float>& h) {
  float* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  m->allocs.push_back(d);
  return d;
}

// round-to-nearest (ties away) to the 10-bit tf32 mantissa, like cvt.rna.tf32.f32
float tf32_rna_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

// w: [cout][cin*ks*ks] (OIHW flattened).  GEMM K order: channel for 1x1, tap*cin + c for 3x3.
bool pack(DetModel* m, PackedConv& pc, const std::vector<float>& w, const std::vector<float>& b, int cout, int cin,
          int ks, int mode) {
  pc.cin = cin;
  pc.cout = cout;
  pc.mode = mode;
  pc.kdim = cin * ks * ks;
  pc.nkb = (pc.kdim + 31) / 32;
  const int kpad = pc.nkb * 32;
  pc.npad_total = (cout + 15) / 16 * 16;
  pc.n_tiles_n = pc.npad_total > 256 ? 2 : 1;
  pc.nt = pc.npad_total / pc.n_tiles_n;
  if (pc.nt % 16 != 0 || cin % 4 != 0) return false;
  std::vector<float> wp((size_t)2 * pc.npad_total * kpad, 0.f), bp(pc.npad_total, 0.f);
  const int taps = ks * ks;
  for (int co = 0; co < cout; ++co) {
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        const float v = w[((size_t)co * cin + ci) * taps + t];
        const float h = tf32_rna_host(v);
        const size_t kk = (size_t)t * cin + ci;
        wp[(size_t)co * kpad + kk] = h;
        wp[((size_t)pc.npad_total + co) * kpad + kk] = v - h;
      }
    bp[co] = b[co];
  }
  pc.wpack = upload(m, wp);
  pc.bias = upload(m, bp);
  if (!pc.wpack || !pc.bias) return false;
  return tc_make_map_2d_f32(&pc.tmW, pc.wpack, (uint64_t)2 * pc.npad_total, (uint64_t)kpad, (uint64_t)kpad,
                            (uint32_t)pc.nt);
}

float* act_alloc(DetModel* m, size_t elems) {
  void* p = nullptr;
  if (cudaMalloc(&p, elems * sizeof(float)) != cudaSuccess) return nullptr;
  m->act_allocs.push_back(p);
  return reinterpret_cast<float*>(p);
}

void det_free_acts(DetModel* m) {
  for (void* p : m->act_allocs) cudaFree(p);
  m->act_allocs.clear();
  m->a_stage.clear();
  m->cap = 0;
}

int det_build_acts(fr_ctx* ctx, int cap) {
  DetModel* m = ctx->det;
  if (m->cap >= cap) return FR_OK;
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  det_free_acts(m);
  bool ok = true;
  auto A = [&](size_t per_img) {
    float* p = act_alloc(m, per_img * cap);
    if (!p) ok = false;
    return p;
  };
  m->a_stem = A((size_t)16 * 320 * 320);
  m->a_b0 = A((size_t)16 * 320 * 320);
  int hw = 320;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kStages[s][0]; ++b) {
      if (b == 0) hw /= 2;
      m->a_stage.push_back(A((size_t)kStages[s][1] * hw * hw));
    }
  const int fh[3] = {80, 40, 20};
  for (int i = 0; i < 3; ++i) {
    const size_t px = (size_t)fh[i] * fh[i];
    m->lat[i] = A(16 * px);
    m->inter[i] = A(16 * px);
    m->pout[i] = A(16 * px);
    m->tw0[i] = A(64 * px);
    m->tw1[i] = A(64 * px);
    m->score[i] = A(2 * px);
    m->bbox[i] = A(8 * px);
    m->kps[i] = A(20 * px);
  }
  if (!ok) {
    det_free_acts(m);
    return fr_fail(ctx, FR_ERR_CUDA, "det activation allocation failed");
  }
  m->cap = cap;
  return FR_OK;
}

struct LayerIO {
  const float* in = nullptr;
  float* out = nullptr;
  int hin = 0, stride = 1, relu = 0, accumulate = 0;
  const float* add_up = nullptr;
  int head = -1;
};

int launch_layer(fr_ctx* ctx, const PackedConv& pc, const LayerIO& io, int n) {
  DetModel* m = ctx->det;
  SepParams p;
  memset(&p, 0, sizeof(p));
  p.in = io.in;
  p.out = io.out;
  p.dw_w = pc.dw_w;
  p.dw_b = pc.dw_b;
  p.bias = pc.bias;
  p.add_up = io.add_up;
  p.cin = pc.cin;
  p.cout = pc.cout;
  p.kdim = pc.kdim;
  p.hin = p.win = io.hin;
  p.hout = p.wout = io.hin / io.stride;
  p.stride = io.stride;
  p.mode = pc.mode;
  p.relu = io.relu;
  p.accumulate = io.accumulate;
  p.epi = io.head >= 0 ? EPI_HEAD : EPI_STD;
  if (io.head >= 0) {
    p.score = m->score[io.head];
    p.bbox = m->bbox[io.head];
    p.kps = m->kps[io.head];
  }
  p.nkb = pc.nkb;
  p.nt = pc.nt;
  p.n_tiles_n = pc.n_tiles_n;
  p.npad_total = pc.npad_total;
  p.total_px = n * p.hout * p.wout;
  if (p.wout % 16 == 0 && p.hout % 8 == 0) {
    p.tiles_x = p.wout / 16;
    p.tiles_y = p.hout / 8;
    p.num_m_tiles = n * p.tiles_x * p.tiles_y;
  } else {
    p.num_m_tiles = ceil_div(p.total_px, TM);
  }
  const int stage_bytes = 2 * A_BYTES + 2 * pc.nt * 128;
  const int budget = 227 * 1024 - 1024 - 512;
  p.stages = std::min(MAX_STAGES, budget / stage_bytes);
  if (p.stages < 2) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd layer does not fit shared memory");
  static const int nbig_env = getenv("FR_SCRFD_NBIG") ? atoi(getenv("FR_SCRFD_NBIG")) : 2;
  p.nbig = (nbig_env == 1 || pc.kdim <= 16) ? 1 : 2;
  const int slot = (p.nbig + 1) * pc.nt;
  if (slot > 512) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd layer does not fit tensor memory");
  p.acc_stages = 2 * slot <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * slot) cols *= 2;
  p.tmem_cols = cols;
  p.err_flag = m->err_flag;
  const int smem = p.stages * stage_bytes + 1024 + 512;
  const int total_tiles = p.num_m_tiles * p.n_tiles_n;
  const int grid = std::min(total_tiles, m->num_sms);
  sep_gemm_kernel<<<grid, THREADS, smem, ctx->stream>>>(pc.tmW, p);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

}  // namespace

int det_model_create(fr_ctx* ctx, const fr_weights* w) {
  if (!w || w->model != FR_MODEL_DET) return fr_fail(ctx, FR_ERR_MODEL, "det weights missing");
  std::unique_ptr<DetModel> m(new DetModel());
  bool ok = true;
  auto dense = [&](const std::string& name) {   // conv (3x3 or 1x1) as a [cout][cin*k*k] GEMM
    const fr_tensor& tw = w->at(name + ".w");
    const int cout = (int)tw.dims[0], cin = (int)tw.dims[1], k = (int)tw.dims[2];
    ok = ok && pack(m.get(), m->conv[name], tw.data, w->at(name + ".b").data, cout, cin, k, k == 3 ? LD_IM2COL : LD_PW);
  };
  auto dwsep = [&](const std::string& name) {   // dw 3x3 + ReLU evaluated inside the 1x1's operand gather
    const fr_tensor& pw = w->at(name + ".pw.w");
    const int cout = (int)pw.dims[0], cin = (int)pw.dims[1];
    PackedConv& pc = m->conv[name];
    ok = ok && pack(m.get(), pc, pw.data, w->at(name + ".pw.b").data, cout, cin, 1, LD_DW);
    const std::vector<float>& dw = w->at(name + ".dw.w").data;   // [cin][1][3][3] -> [9][cin]
    std::vector<float> dwt((size_t)9 * cin);
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) dwt[(size_t)t * cin + c] = dw[(size_t)c * 9 + t];
    pc.dw_w = upload(m.get(), dwt);
    pc.dw_b = upload(m.get(), w->at(name + ".dw.b").data);
    ok = ok && pc.dw_w && pc.dw_b;
  };
  {
    const fr_tensor& tw = w->at("stem.w");   // [16][3][3][3] -> [27][16]
    std::vector<float> sw(27 * 16);
    for (int co = 0; co < 16; ++co)
      for (int k = 0; k < 27; ++k) sw[k * 16 + co] = tw.data[co * 27 + k];
    m->stem_w = upload(m.get(), sw);
    m->stem_b = upload(m.get(), w->at("stem.b").data);
    ok = ok && m->stem_w && m->stem_b;
  }
  dwsep("b0");
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kStages[s][0]; ++b) dwsep("s" + std::to_string(s) + "." + std::to_string(b));
  for (int i = 0; i < 3; ++i) dense("lat" + std::to_string(i));
  for (int i = 0; i < 3; ++i) dense("fpn" + std::to_string(i));
  for (int i = 0; i < 2; ++i) dense("down" + std::to_string(i));
  for (int i = 0; i < 2; ++i) dense("pafpn" + std::to_string(i));
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    dwsep(h + ".t0");
    dwsep(h + ".t1");
    // fuse the three head convs of the stride into one 64 -> 30 conv (cls 2 | reg 8 | kps 20)
    std::vector<float> fw, fb;
    for (const char* part : {".cls", ".reg", ".kps"}) {
      const fr_tensor& tw = w->at(h + part + ".w");
      const fr_tensor& tb = w->at(h + part + ".b");
      fw.insert(fw.end(), tw.data.begin(), tw.data.end());
      fb.insert(fb.end(), tb.data.begin(), tb.data.end());
    }
    ok = ok && pack(m.get(), m->conv[h + ".out"], fw, fb, 30, 64, 3, LD_IM2COL);
  }
  if (ok) {
    ok = cudaMalloc(&m->err_flag, sizeof(int)) == cudaSuccess;
    if (ok) {
      cudaMemset(m->err_flag, 0, sizeof(int));
      m->allocs.push_back(m->err_flag);
    }
  }
  if (ok) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess) m->num_sms = prop.multiProcessorCount;
    ok = cudaFuncSetAttribute(sep_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
  }
  if (!ok) {
    for (void* p : m->allocs) cudaFree(p);
    return fr_fail(ctx, FR_ERR_CUDA, "det weight upload / tensor map creation failed");
  }
  ctx->det = m.release();
  return FR_OK;
}

void det_model_destroy(fr_ctx* ctx) {
  DetModel* m = ctx->det;
  if (!m) return;
  det_free_acts(m);
  for (void* p : m->allocs) cudaFree(p);
  delete m;
  ctx->det = nullptr;
}

int det_forward(fr_ctx* ctx, const __nv_bfloat16* d_in_chw, int n, HeadPtrs* heads) {
  DetModel* m = ctx->det;
  if (!m) return fr_fail(ctx, FR_ERR_NOT_LOADED, "Model not loaded!");
  if (n <= 0) return FR_OK;
  int cap = 1;
  while (cap < n) cap *= 2;
  FR_CHECK(det_build_acts(ctx, cap));
  // stem: 3x3 s2, 3 -> 16, ReLU (bf16 planar input from K1) -> fp32 NHWC
  {
    const size_t total = (size_t)n * (DET / 2) * (DET / 2);
    stem_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_in_chw, m->a_stem, m->stem_w,
                                                                                m->stem_b, n);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
  }
  auto dwsep = [&](const std::string& name, const float* in, float* out, int hin, int stride) -> int {
    LayerIO io;
    io.in = in; io.out = out; io.hin = hin; io.stride = stride; io.relu = 1;
    return launch_layer(ctx, m->conv.at(name), io, n);
  };
  FR_CHECK(dwsep("b0", m->a_stem, m->a_b0, 320, 1));
  const float* cur = m->a_b0;
  int hw = 320, bi = 0;
  const float* feats[3] = {nullptr, nullptr, nullptr};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < kStages[s][0]; ++b, ++bi) {
      const int stride = b == 0 ? 2 : 1;
      FR_CHECK(dwsep("s" + std::to_string(s) + "." + std::to_string(b), cur, m->a_stage[bi], hw, stride));
      hw /= stride;
      cur = m->a_stage[bi];
    }
    if (s >= 1) feats[s - 1] = cur;
  }
  const int fh[3] = {80, 40, 20};
  // laterals (1x1, no activation) with the top-down nearest-2x add fused into the epilogue
  for (int i = 2; i >= 0; --i) {
    LayerIO io;
    io.in = feats[i]; io.out = m->lat[i]; io.hin = fh[i];
    io.add_up = i < 2 ? m->lat[i + 1] : nullptr;
    FR_CHECK(launch_layer(ctx, m->conv.at("lat" + std::to_string(i)), io, n));
  }
  auto conv3 = [&](const std::string& name, const float* in, float* out, int hin, int stride,
                   int accumulate) -> int {
    LayerIO io;
    io.in = in; io.out = out; io.hin = hin; io.stride = stride; io.accumulate = accumulate;
    return launch_layer(ctx, m->conv.at(name), io, n);
  };
  for (int i = 0; i < 3; ++i) FR_CHECK(conv3("fpn" + std::to_string(i), m->lat[i], m->inter[i], fh[i], 1, 0));
  for (int i = 0; i < 2; ++i)
    FR_CHECK(conv3("down" + std::to_string(i), m->inter[i], m->inter[i + 1], fh[i], 2, 1));
  const float* outs[3] = {m->inter[0], m->pout[1], m->pout[2]};
  for (int i = 1; i < 3; ++i)
    FR_CHECK(conv3("pafpn" + std::to_string(i - 1), m->inter[i], m->pout[i], fh[i], 1, 0));
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    FR_CHECK(dwsep(h + ".t0", outs[i], m->tw0[i], fh[i], 1));
    FR_CHECK(dwsep(h + ".t1", m->tw0[i], m->tw1[i], fh[i], 1));
    LayerIO io;
    io.in = m->tw1[i]; io.hin = fh[i]; io.head = i;
    FR_CHECK(launch_layer(ctx, m->conv.at(h + ".out"), io, n));
    heads->score[i] = m->score[i];
    heads->bbox[i] = m->bbox[i];
    heads->kps[i] = m->kps[i];
  }
  return FR_OK;
}

// Test hook: copy one intermediate activation (fp32 NHWC) of the last det_forward to the host.
// tap: 0 stem, 1 b0, 2..14 backbone blocks, 15..17 lat, 18..20 inter, 21..22 pout[1..2],
// 23..25 head tower 0, 26..28 head tower 1.
int det_tap(fr_ctx* ctx, int tap, int n, float* h_out, size_t out_elems) {
  DetModel* m = ctx->det;
  if (!m || m->cap < n) return fr_fail(ctx, FR_ERR_NOT_LOADED, "no detector activations");
  const float* src = nullptr;
  if (tap == 0) src = m->a_stem;
  else if (tap == 1) src = m->a_b0;
  else if (tap >= 2 && tap < 15) src = m->a_stage[tap - 2];
  else if (tap < 18) src = m->lat[tap - 15];
  else if (tap < 21) src = m->inter[tap - 18];
  else if (tap < 23) src = m->pout[tap - 20];
  else if (tap < 26) src = m->tw0[tap - 23];
  else if (tap < 29) src = m->tw1[tap - 26];
  if (!src) return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad tap");
  FR_CUDA_OK(ctx, cudaMemcpyAsync(h_out, src, out_elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return FR_OK;
}
