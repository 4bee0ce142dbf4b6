// K9: 1:N cosine search with fused top-k (north_star extension; the reference only has the
// 1:1 FaceRecognizer::compareFaces, src/face_recognizer.cpp:320-334, whose batched
// generalisation this is: S = Q . G^T on L2-normalised rows, top-k per query, ties -> lower
// index; mapped score (S+1)/2 and the 0.6 rule are applied by the caller).
//
// One tcgen05 kernel does GEMM and selection; the 4096 x N score matrix never exists:
//   * a CTA owns one tile of 128 queries (A operand: 128 x 512 bf16 = 128 KB, loaded once by
//     TMA and kept resident in shared memory) and sweeps a contiguous range of gallery tiles
//   * gallery tiles (256 rows x 64 K-columns, SWIZZLE_128B) stream through a 3-stage TMA ring
//   * tcgen05.mma M=128 N=256 K=16, fp32 accumulators double-buffered in TMEM (2 x 256 cols)
//   * epilogue warps read the accumulator with tcgen05.ld; each thread owns one query row and
//     keeps its running top-k in registers across all tiles of the sweep
// Partial lists of the splits (and, multi-GPU, of the ranks after the NCCL all-gather) are
// merged by topk_merge_kernel with the (score desc, global index asc) order, so the result
// does not depend on the number of splits or ranks.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp8.h>

#include "common.h"
#include "tc_gemm.cuh"

bool tc_make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                    uint64_t pitch_elems, uint32_t box_rows);
bool tc_make_map_2d_u8(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t pitch_bytes, uint32_t box_rows);

namespace {

typedef __nv_bfloat16 bf16;
constexpr int DIM = FR_FEAT_DIM;        // 512
constexpr int KB = DIM / tc::BK;        // 8 K blocks
constexpr int GN = 256;                 // gallery rows per tile
constexpr int TOPK = 16;                // list capacity (k <= 16)
constexpr int G_STAGES = 3;
constexpr int G_B_BYTES = GN * tc::BK * 2;             // 32 KB
constexpr int G_A_BYTES = KB * tc::A_TILE_BYTES;       // 128 KB
constexpr int G_SMEM = G_A_BYTES + G_STAGES * G_B_BYTES + 256 + 1024;
// e4m3 coarse pass (kind::f8f6f4, K = 32 per MMA): a 128-byte swizzle row holds 128 elements, so the
// 512-d rows are 4 K blocks; operand tiles have the same byte sizes, the resident query tile is half.
// Values are stored as e4m3(x * 2^6): unit-norm 512-d rows have |x| ~ 0.044, which would sit on e4m3's
// subnormal edge (min normal 2^-6); the power-of-two scale is undone exactly on the accumulator.
constexpr int KB8 = DIM / 128;
constexpr float FP8_SCALE = 64.f;
constexpr float FP8_INV_SCALE2 = 1.f / (FP8_SCALE * FP8_SCALE);

struct GParams {
  int n_rows;           // gallery rows in this shard
  int nq;               // valid queries
  int num_m_tiles;
  int n_tiles;          // ceil(n_rows / GN)
  int tiles_per_split;
  int nq_pad;
  float* out_s;         // [splits][nq_pad][TOPK]
  int* out_i;           // [splits][nq_pad][TOPK]
  int* err_flag;
};

// Packed candidate record for the rank exchange: low word = fp32 score bits, high word = global row
// index (0xffffffff = empty slot).  One 8-byte record per candidate makes the multi-GPU exchange a
// single all-gather.
__device__ __forceinline__ unsigned long long pack_rec(float s, long long i) {
  return (unsigned long long)__float_as_uint(s) | ((unsigned long long)(i < 0 ? 0xffffffffu : (uint32_t)i) << 32);
}

// kind::f8f6f4 instruction descriptor: e4m3 x e4m3 -> fp32, both operands K-major (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_e4m3(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// FP8 = false: bf16 operands (kind::f16, 8 K blocks of 64); FP8 = true: e4m3 operands (kind::f8f6f4, 4 K
// blocks of 128), scores rescaled by 2^-12 on the way out.
template <bool FP8>
__global__ void __launch_bounds__(tc::NUM_THREADS, 1)
gallery_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ GParams p) {
  using namespace tc;
  constexpr int KB = FP8 ? KB8 : ::KB;                 // K blocks per row
  constexpr int KSTEP = FP8 ? 128 : BK;                // elements per K block (128 bytes either way)
  constexpr int G_A_BYTES = KB * tc::A_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + G_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + G_STAGES * G_B_BYTES);
  uint64_t* empty = full + G_STAGES;
  uint64_t* tfull = empty + G_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x % p.num_m_tiles;
  const int split = blockIdx.x / p.num_m_tiles;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.n_tiles);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmG);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(afull, G_A_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * A_TILE_BYTES, &tmQ, afull, kb * KSTEP, m_tile * BM);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t)
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, p.err_flag);
          mbar_expect_tx(&full[stage], G_B_BYTES);
          tma_load_2d(sB + stage * G_B_BYTES, &tmG, &full[stage], kb * KSTEP, t * GN);
          if (++stage == G_STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(BM, GN) : make_idesc(BM, GN);
      mbar_wait(afull, 0, p.err_flag);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase, p.err_flag);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sA + kb * A_TILE_BYTES);
          const uint64_t bdesc = make_smem_desc(sB + stage * G_B_BYTES);
          // four MMAs per K block either way: K = 16 bf16 or K = 32 e4m3 = 32 bytes = +2 descriptor units
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (FP8) mma_e4m3(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            else mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);
          if (++stage == G_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tfull[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = m_tile * BM + q * 32 + lane;
    float ts[TOPK];
    int ti[TOPK];
#pragma unroll
    for (int j = 0; j < TOPK; ++j) { ts[j] = -INFINITY; ti[j] = -1; }
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * GN + ((uint32_t)(q * 32) << 16);
      const int col0 = t * GN;
#pragma unroll 1
      for (int c = 0; c < GN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr0 + c * 32, v);
        const int cb = col0 + c * 32;
        // Fast path: one compare per score against the current k-th best builds a candidate
        // bit mask; almost always it is empty.  (Written as a branch on the mask so that the
        // compiler cannot if-convert the 16-step insertion into every element's path.)
        const float thr = ts[TOPK - 1];
        uint32_t mask = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(v[i]) > thr) ? (1u << i) : 0u;
        while (mask) {
          const int i = __ffs(mask) - 1;
          mask &= mask - 1;
          // v[i] with a runtime i: 5-level select tree over the registers
          uint32_t a16[16], a8[8], a4[4], a2[2];
#pragma unroll
          for (int j = 0; j < 16; ++j) a16[j] = (i & 16) ? v[16 + j] : v[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) a8[j] = (i & 8) ? a16[8 + j] : a16[j];
#pragma unroll
          for (int j = 0; j < 4; ++j) a4[j] = (i & 4) ? a8[4 + j] : a8[j];
#pragma unroll
          for (int j = 0; j < 2; ++j) a2[j] = (i & 2) ? a4[2 + j] : a4[j];
          float cs = __uint_as_float((i & 1) ? a2[1] : a2[0]);
          int ci = cb + i;
          if (cs > ts[TOPK - 1] && ci < p.n_rows) {   // re-check: the threshold rises as we insert
            bool ins = false;   // once inserted, everything below shifts down by one
#pragma unroll
            for (int j = 0; j < TOPK; ++j)
              if (ins || cs > ts[j]) {   // increasing-index sweep: strict '>' keeps ties stable
                const float fs = ts[j]; ts[j] = cs; cs = fs;
                const int fi = ti[j]; ti[j] = ci; ci = fi;
                ins = true;
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (row < p.nq) {
      float* os = p.out_s + ((size_t)split * p.nq_pad + row) * TOPK;
      int* oi = p.out_i + ((size_t)split * p.nq_pad + row) * TOPK;
#pragma unroll
      for (int j = 0; j < TOPK; ++j) { os[j] = FP8 ? ts[j] * FP8_INV_SCALE2 : ts[j]; oi[j] = ti[j]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// fp32 / bf16 rows -> e4m3(x * 2^6), saturating
__global__ void rows_to_e4m3_kernel(const float* __restrict__ in_f32, const bf16* __restrict__ in_bf16,
                                    uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float v = in_f32 ? __bfloat162float(__float2bfloat16_rn(in_f32[i])) : __bfloat162float(in_bf16[i]);
    out[i] = (uint8_t)__nv_cvt_float_to_fp8(v * FP8_SCALE, __NV_SATFINITE, __NV_E4M3);
  }
}

// Reduce `parts` per-split candidate lists to `groups` lists (group g merges parts [g*ppg, (g+1)*ppg)) by
// coarse score, keeping local indices: bounds the re-rank's candidate set at groups * TOPK when a small
// query batch is spread over many splits.
__global__ void coarse_group_merge_kernel(const float* __restrict__ s_in, const int* __restrict__ i_in, int parts,
                                          int ppg, int nq, int part_stride, float* __restrict__ s_out,
                                          int* __restrict__ i_out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (qi >= nq) return;
  float ts[TOPK];
  int ti[TOPK];
#pragma unroll
  for (int j = 0; j < TOPK; ++j) { ts[j] = -INFINITY; ti[j] = -1; }
  for (int pt = g * ppg; pt < min((g + 1) * ppg, parts); ++pt)
    for (int c = 0; c < TOPK; ++c) {
      const size_t o = ((size_t)pt * part_stride + qi) * TOPK + c;
      float cs = s_in[o];
      int ci = i_in[o];
      if (ci < 0) continue;
      bool ins = false;
#pragma unroll
      for (int j = 0; j < TOPK; ++j) {
        const bool better = ins || ti[j] < 0 || cs > ts[j] || (cs == ts[j] && ci < ti[j]);
        if (better) {
          const float fs = ts[j]; ts[j] = cs; cs = fs;
          const int fi = ti[j]; ti[j] = ci; ci = fi;
          ins = true;
        }
      }
    }
  const size_t o = ((size_t)g * part_stride + qi) * TOPK;
#pragma unroll
  for (int j = 0; j < TOPK; ++j) { s_out[o + j] = ts[j]; i_out[o + j] = ti[j]; }
}

// Exact re-rank of the coarse candidates: one warp per query.  The candidate set is the union of
// `parts` (<= RERANK_MAX_PARTS) top-16 lists of the e4m3 pass (local row indices, -1 = empty).  Each
// candidate's score is recomputed from the bf16 row and the bf16 query with fp32 accumulation, then the
// top `kout` are selected by (score desc, global index asc).
constexpr int RERANK_MAX_PARTS = 8;      // 128 candidates = 4 per lane
__global__ void __launch_bounds__(256)
rerank_kernel(const bf16* __restrict__ q, const bf16* __restrict__ rows, const int* __restrict__ cand, int parts,
              int nq, int part_stride, int kout, long long base, float* __restrict__ s_out,
              long long* __restrict__ i_out, unsigned long long* __restrict__ rec_out) {
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  // this lane's 16 query values (2 x 16-byte loads): dims [lane*8, +8) and [256 + lane*8, +8)
  float qv[16];
  {
    const uint4 a = *reinterpret_cast<const uint4*>(q + (size_t)qi * DIM + lane * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(q + (size_t)qi * DIM + 256 + lane * 8);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      qv[2 * j] = __uint_as_float(w[j] << 16);
      qv[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  const int n_cand = parts * TOPK;
  float my_s[4];
  long long my_i[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { my_s[j] = -INFINITY; my_i[j] = -1; }
  for (int c = 0; c < n_cand; ++c) {
    const int pt = c / TOPK, e = c - pt * TOPK;
    const int l = cand[((size_t)pt * part_stride + qi) * TOPK + e];
    if (l < 0) continue;                    // warp-uniform
    const bf16* r = rows + (size_t)l * DIM;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(r + lane * 8));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(r + 256 + lane * 8));
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc = fmaf(qv[2 * j], __uint_as_float(w[j] << 16), acc);
      acc = fmaf(qv[2 * j + 1], __uint_as_float(w[j] & 0xffff0000u), acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((c & 31) == lane) {
      const int slot = c >> 5;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == slot) { my_s[j] = acc; my_i[j] = base + l; }
    }
  }
  // kout rounds of warp arg-max over (score desc, index asc)
  for (int k = 0; k < kout; ++k) {
    float bs = -INFINITY;
    long long bi = -1;
    int bj = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (my_i[j] >= 0 && (bi < 0 || my_s[j] > bs || (my_s[j] == bs && my_i[j] < bi))) { bs = my_s[j]; bi = my_i[j]; bj = j; }
    float ws = bs;
    long long wi = bi;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws, off);
      const long long oi = __shfl_xor_sync(0xffffffffu, wi, off);
      if (oi >= 0 && (wi < 0 || os > ws || (os == ws && oi < wi))) { ws = os; wi = oi; }
    }
    if (wi >= 0 && wi == bi) {            // the winner's owner retires it (candidate indices are unique)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == bj) my_i[j] = -1;
    }
    if (lane == 0) {
      const float fs = wi < 0 ? -INFINITY : ws;
      if (rec_out) rec_out[(size_t)qi * kout + k] = pack_rec(fs, wi);
      else { s_out[(size_t)qi * kout + k] = fs; i_out[(size_t)qi * kout + k] = wi; }
    }
  }
}

// Merge `parts` sorted lists of `kin` candidates per query into the top `kout`
// (score desc, global index asc).  idx32 (local, + base) or idx64 (already global) input.
// rec_in (packed lists) replaces s_in / i32_in / i64_in when non-null; rec_out replaces s_out / i_out.
__global__ void topk_merge_kernel(const float* __restrict__ s_in, const int* __restrict__ i32_in,
                                  const long long* __restrict__ i64_in, int parts, int nq, int part_stride,
                                  int kin, int kout, long long base, float* __restrict__ s_out,
                                  long long* __restrict__ i_out,
                                  const unsigned long long* __restrict__ rec_in = nullptr,
                                  unsigned long long* __restrict__ rec_out = nullptr) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  float ts[TOPK];
  long long ti[TOPK];
#pragma unroll
  for (int j = 0; j < TOPK; ++j) { ts[j] = -INFINITY; ti[j] = -1; }
  for (int pt = 0; pt < parts; ++pt)
    for (int c = 0; c < kin; ++c) {
      const size_t o = ((size_t)pt * part_stride + qi) * kin + c;
      float cs;
      long long ci;
      if (rec_in) {
        const unsigned long long r = rec_in[o];
        cs = __uint_as_float((uint32_t)r);
        const uint32_t hi = (uint32_t)(r >> 32);
        ci = hi == 0xffffffffu ? -1 : (long long)hi;
      } else {
        cs = s_in[o];
        if (i64_in) ci = i64_in[o];
        else { const int l = i32_in[o]; ci = l < 0 ? -1 : base + l; }
      }
      if (ci < 0) continue;
      bool ins = false;
#pragma unroll
      for (int j = 0; j < TOPK; ++j) {
        const bool better = ins || ti[j] < 0 || cs > ts[j] || (cs == ts[j] && ci < ti[j]);
        if (better) {
          const float fs = ts[j]; ts[j] = cs; cs = fs;
          const long long fi = ti[j]; ti[j] = ci; ci = fi;
          ins = true;
        }
      }
    }
  for (int j = 0; j < kout; ++j) {
    if (rec_out) {
      rec_out[(size_t)qi * kout + j] = pack_rec(ts[j], ti[j]);
    } else {
      s_out[(size_t)qi * kout + j] = ts[j];
      i_out[(size_t)qi * kout + j] = ti[j];
    }
  }
}

__global__ void rows_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// One warp per row: 512 hashed values in (-1,1), L2-normalised, stored bf16.
__global__ void fill_synthetic_kernel(bf16* __restrict__ g, long long first, long long n, uint64_t seed,
                                      long long global_base) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const uint64_t rowkey = mix64(seed ^ (uint64_t)(global_base + first + r) * 0xD1B54A32D192ED03ull);
  float v[DIM / 32];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < DIM / 32; ++j) {
    const uint64_t h = mix64(rowkey + (uint64_t)(j * 32 + lane));
    v[j] = (float)(h >> 40) * (2.0f / 16777216.0f) - 1.0f;
    ss = fmaf(v[j], v[j], ss);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float inv = rsqrtf(ss);
#pragma unroll
  for (int j = 0; j < DIM / 32; ++j) g[(size_t)(first + r) * DIM + j * 32 + lane] = __float2bfloat16_rn(v[j] * inv);
}

__global__ void bf16_rows_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

}  // namespace

struct fr_gallery {
  fr_ctx* ctx = nullptr;
  bf16* rows = nullptr;        // bf16 rows: HBM, or mapped pinned host memory (FR_GALLERY_BF16_ON_HOST)
  uint8_t* rows8 = nullptr;    // e4m3(x * 64) rows for the coarse pass (FR_GALLERY_FP8), HBM
  void* rows_host = nullptr;   // the host allocation behind `rows` in BF16_ON_HOST mode
  int flags = 0;
  int64_t cap = 0, size = 0, base = 0;
  DevBuf q_bf16, q_f32, q_e4m3, part_s, part_i, grp_s, grp_i, out_s, out_i, rec_local, rec_all;
  int* err_flag = nullptr;
};

namespace {
struct GGuard {
  fr_ctx* c;
  explicit GGuard(fr_ctx* ctx) : c(ctx) { c->mu.lock(); cudaSetDevice(c->device); }
  ~GGuard() { c->mu.unlock(); }
};
}  // namespace

extern "C" {

int fr_gallery_create_ex(fr_ctx* ctx, fr_gallery** out, int64_t capacity_rows, int64_t index_base, int flags) {
  if (!ctx || !out || capacity_rows <= 0) return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad gallery arguments");
  if ((flags & ~(FR_GALLERY_FP8 | FR_GALLERY_BF16_ON_HOST)) || ((flags & FR_GALLERY_BF16_ON_HOST) && !(flags & FR_GALLERY_FP8)))
    return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad gallery flags (BF16_ON_HOST needs FP8)");
  GGuard g(ctx);
  std::unique_ptr<fr_gallery> G(new fr_gallery());
  G->ctx = ctx;
  G->cap = capacity_rows;
  G->base = index_base;
  G->flags = flags;
  // round the allocation up to a whole tile so TMA boxes never leave the allocation
  const size_t alloc_rows = ((size_t)capacity_rows + GN - 1) / GN * GN;
  bool ok = cudaMalloc(&G->err_flag, 4) == cudaSuccess;
  if (flags & FR_GALLERY_BF16_ON_HOST) {
    // capacity mode: only the e4m3 rows live in HBM (512 B / row); the bf16 rows the re-rank reads sit in
    // mapped pinned host memory (unified addressing: the kernels dereference the same pointer)
    ok = ok && cudaHostAlloc(&G->rows_host, alloc_rows * DIM * 2, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    if (ok) {
      memset(G->rows_host, 0, alloc_rows * DIM * 2);
      void* dp = nullptr;
      ok = cudaHostGetDevicePointer(&dp, G->rows_host, 0) == cudaSuccess;
      G->rows = reinterpret_cast<bf16*>(dp);
    }
  } else {
    ok = ok && cudaMalloc(&G->rows, alloc_rows * DIM * 2) == cudaSuccess;
    if (ok) cudaMemsetAsync(G->rows, 0, alloc_rows * DIM * 2, ctx->stream);
  }
  if (flags & FR_GALLERY_FP8) {
    ok = ok && cudaMalloc(&G->rows8, alloc_rows * DIM) == cudaSuccess;
    if (ok) cudaMemsetAsync(G->rows8, 0, alloc_rows * DIM, ctx->stream);
  }
  if (!ok) {
    if (G->rows_host) cudaFreeHost(G->rows_host); else if (G->rows) cudaFree(G->rows);
    if (G->rows8) cudaFree(G->rows8);
    if (G->err_flag) cudaFree(G->err_flag);
    return fr_fail(ctx, FR_ERR_CUDA, "gallery allocation failed");
  }
  cudaMemsetAsync(G->err_flag, 0, 4, ctx->stream);
  *out = G.release();
  return FR_OK;
}

int fr_gallery_create(fr_ctx* ctx, fr_gallery** out, int64_t capacity_rows, int64_t index_base) {
  return fr_gallery_create_ex(ctx, out, capacity_rows, index_base, FR_GALLERY_BF16);
}

void fr_gallery_destroy(fr_gallery* g) {
  if (!g) return;
  {
    GGuard gg(g->ctx);
    cudaStreamSynchronize(g->ctx->stream);
    if (g->rows_host) cudaFreeHost(g->rows_host); else cudaFree(g->rows);
    if (g->rows8) cudaFree(g->rows8);
    g->q_e4m3.release(); g->grp_s.release(); g->grp_i.release();
    cudaFree(g->err_flag);
    g->q_bf16.release(); g->q_f32.release(); g->part_s.release(); g->part_i.release();
    g->out_s.release(); g->out_i.release(); g->rec_local.release(); g->rec_all.release();
  }
  delete g;
}

int64_t fr_gallery_size(const fr_gallery* g) { return g ? g->size : 0; }

int fr_gallery_add(fr_gallery* g, const float* rows, int64_t n, int memspace) {
  if (!g || !rows || n <= 0) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  if (g->size + n > g->cap) return fr_fail(ctx, FR_ERR_CAPACITY, "gallery full");
  const size_t elems = (size_t)n * DIM;
  const float* d_rows = rows;
  if (memspace != FR_MEM_DEVICE) {
    if (!g->q_f32.reserve(elems * 4)) return fr_fail(ctx, FR_ERR_CUDA, "gallery staging allocation failed");
    FR_CUDA_OK(ctx, cudaMemcpyAsync(g->q_f32.p, rows, elems * 4, cudaMemcpyHostToDevice, ctx->stream));
    d_rows = g->q_f32.as<float>();
  }
  rows_to_bf16_kernel<<<148 * 4, 256, 0, ctx->stream>>>(d_rows, g->rows + (size_t)g->size * DIM, elems);
  ctx->launches++;
  if (g->rows8) {   // e4m3 mirror of the bf16-rounded values (what a save -> load round trip would quantise)
    rows_to_e4m3_kernel<<<148 * 4, 256, 0, ctx->stream>>>(d_rows, nullptr, g->rows8 + (size_t)g->size * DIM, elems);
    ctx->launches++;
  }
  FR_CUDA_OK(ctx, cudaGetLastError());
  if (memspace != FR_MEM_DEVICE) FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  g->size += n;
  return FR_OK;
}

int fr_gallery_fill_synthetic(fr_gallery* g, int64_t n, uint64_t seed) {
  if (!g || n <= 0) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  if (g->size + n > g->cap) return fr_fail(ctx, FR_ERR_CAPACITY, "gallery full");
  const unsigned blocks = (unsigned)((n + 7) / 8);
  fill_synthetic_kernel<<<blocks, 256, 0, ctx->stream>>>(g->rows, g->size, n, seed, g->base);
  ctx->launches++;
  if (g->rows8) {
    rows_to_e4m3_kernel<<<148 * 8, 256, 0, ctx->stream>>>(nullptr, g->rows + (size_t)g->size * DIM,
                                                          g->rows8 + (size_t)g->size * DIM, (size_t)n * DIM);
    ctx->launches++;
  }
  FR_CUDA_OK(ctx, cudaGetLastError());
  g->size += n;
  return FR_OK;
}

int fr_gallery_get_rows(fr_gallery* g, int64_t first, int64_t n, float* out_host) {
  if (!g || !out_host || first < 0 || n <= 0 || first + n > g->size) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  const size_t elems = (size_t)n * DIM;
  if (!g->q_f32.reserve(elems * 4)) return fr_fail(ctx, FR_ERR_CUDA, "gallery staging allocation failed");
  bf16_rows_to_f32_kernel<<<148 * 4, 256, 0, ctx->stream>>>(g->rows + (size_t)first * DIM, g->q_f32.as<float>(), elems);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaMemcpyAsync(out_host, g->q_f32.p, elems * 4, cudaMemcpyDeviceToHost, ctx->stream));
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return FR_OK;
}

// ---- persistence / enrolment.  File = 32-byte header + size x 512 bf16 rows:
//   char magic[8] = "FRGAL001"; int64 rows; int64 index_base; int32 dim; int32 dtype (1 = bf16)
// The rows are stored exactly as they sit in HBM, so save -> load round-trips bit for bit and a
// shard written by rank r of an N-rank job can be loaded by any rank of any job.
namespace {
struct GalHeader {
  char magic[8];
  int64_t rows, index_base;
  int32_t dim, dtype;
};
static_assert(sizeof(GalHeader) == 32, "gallery file header is 32 bytes");
constexpr size_t kIoChunkRows = 1 << 16;   // 64 MiB staging chunks
}  // namespace

int fr_gallery_save(fr_gallery* g, const char* path) {
  if (!g || !path) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  FILE* f = fopen(path, "wb");
  if (!f) return fr_fail(ctx, FR_ERR_IO, std::string("cannot open for writing: ") + path);
  GalHeader h;
  memcpy(h.magic, "FRGAL001", 8);
  h.rows = g->size;
  h.index_base = g->base;
  h.dim = DIM;
  h.dtype = 1;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  std::vector<uint16_t> buf(std::min<size_t>(kIoChunkRows, (size_t)std::max<int64_t>(g->size, 1)) * DIM);
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int64_t r = 0; ok && r < g->size; r += kIoChunkRows) {
    const size_t n = (size_t)std::min<int64_t>(kIoChunkRows, g->size - r);
    if (cudaMemcpy(buf.data(), g->rows + (size_t)r * DIM, n * DIM * 2, cudaMemcpyDefault) != cudaSuccess) ok = false;
    ok = ok && fwrite(buf.data(), 2, n * DIM, f) == n * DIM;
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? FR_OK : fr_fail(ctx, FR_ERR_IO, std::string("short write: ") + path);
}

// Appends the rows of a saved shard to `g` (which keeps its own index_base); *file_index_base,
// if not null, receives the base recorded in the file.
int fr_gallery_load(fr_gallery* g, const char* path, int64_t* file_index_base) {
  if (!g || !path) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  FILE* f = fopen(path, "rb");
  if (!f) return fr_fail(ctx, FR_ERR_IO, std::string("cannot open: ") + path);
  GalHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "FRGAL001", 8) != 0 || h.dim != DIM || h.dtype != 1 ||
      h.rows < 0) {
    fclose(f);
    return fr_fail(ctx, FR_ERR_IO, std::string("not a gallery file: ") + path);
  }
  if (g->size + h.rows > g->cap) {
    fclose(f);
    return fr_fail(ctx, FR_ERR_CAPACITY, "gallery full");
  }
  std::vector<uint16_t> buf(std::min<size_t>(kIoChunkRows, (size_t)std::max<int64_t>(h.rows, 1)) * DIM);
  FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  bool ok = true;
  for (int64_t r = 0; ok && r < h.rows; r += kIoChunkRows) {
    const size_t n = (size_t)std::min<int64_t>(kIoChunkRows, h.rows - r);
    ok = fread(buf.data(), 2, n * DIM, f) == n * DIM;
    if (ok && cudaMemcpy(g->rows + (size_t)(g->size + r) * DIM, buf.data(), n * DIM * 2, cudaMemcpyDefault) !=
                  cudaSuccess)
      ok = false;
  }
  fclose(f);
  if (!ok) return fr_fail(ctx, FR_ERR_IO, std::string("short read: ") + path);
  if (g->rows8 && h.rows > 0) {
    rows_to_e4m3_kernel<<<148 * 8, 256, 0, ctx->stream>>>(nullptr, g->rows + (size_t)g->size * DIM,
                                                          g->rows8 + (size_t)g->size * DIM, (size_t)h.rows * DIM);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  g->size += h.rows;
  if (file_index_base) *file_index_base = h.index_base;
  return FR_OK;
}

// Removes local row `row` by moving the last row into its place (O(1); the moved row's index
// changes from size-1 to `row`, which the caller's id table must mirror).
int fr_gallery_remove(fr_gallery* g, int64_t row) {
  if (!g || row < 0 || row >= g->size) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  const int64_t last = g->size - 1;
  if (row != last)
    FR_CUDA_OK(ctx, cudaMemcpyAsync(g->rows + (size_t)row * DIM, g->rows + (size_t)last * DIM, DIM * 2,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
  FR_CUDA_OK(ctx, cudaMemsetAsync(g->rows + (size_t)last * DIM, 0, DIM * 2, ctx->stream));
  if (g->rows8) {
    if (row != last)
      FR_CUDA_OK(ctx, cudaMemcpyAsync(g->rows8 + (size_t)row * DIM, g->rows8 + (size_t)last * DIM, DIM,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemsetAsync(g->rows8 + (size_t)last * DIM, 0, DIM, ctx->stream));
  }
  g->size = last;
  return FR_OK;
}

}  // extern "C"

// Local search of one shard, everything enqueued on ctx->stream (caller holds the ctx lock).
// Results: (d_os, d_oi) device arrays, or packed records d_rec (then d_os / d_oi are ignored).
// fp8 = true: coarse pass on the e4m3 rows (kind::f8f6f4, twice the MMA rate, half the bytes), then an exact
// bf16 re-rank of the union of the per-split top-16 lists (up to 128 candidates per query).
static int gallery_search_local(fr_gallery* g, const float* queries, int nq, int k, int memspace, float* d_os,
                                long long* d_oi, unsigned long long* d_rec, bool fp8 = false) {
  fr_ctx* ctx = g->ctx;
  const int num_sms = ctx->num_sms;
  constexpr int G_SMEM8 = KB8 * tc::A_TILE_BYTES + G_STAGES * G_B_BYTES + 256 + 1024;
  if (fp8 && !g->rows8) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "gallery was created without FR_GALLERY_FP8");
  if (!fp8 && (g->flags & FR_GALLERY_BF16_ON_HOST))
    return fr_fail(ctx, FR_ERR_UNSUPPORTED, "bf16 rows live in host memory (FR_GALLERY_BF16_ON_HOST): use fr_gallery_search_fp8");
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, gallery_topk_kernel<false>, G_SMEM));
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, gallery_topk_kernel<true>, G_SMEM8));
  const int num_m_tiles = ceil_div(nq, tc::BM);
  const int nq_pad = num_m_tiles * tc::BM;
  const int n_tiles = (int)((g->size + GN - 1) / GN);
  // Split the gallery sweep so that the grid (num_m_tiles x splits CTAs, one per SM at a time) fills whole waves:
  // 32 query tiles x 4 splits = 128 CTAs leave 20 of 148 SMs idle (0.86); 32 x 9 = 288 CTAs = 1.95 waves (0.97).
  // Smallest split count within 3 % of the best wave efficiency (more splits = more partial lists to merge).
  int splits = 1;
  if (n_tiles > 0) {
    double best = 0.0;
    const int smax = std::min(n_tiles, std::max(24, ceil_div(num_sms, num_m_tiles)));   // small batches: one CTA per SM
    for (int s2 = 1; s2 <= smax; ++s2) {
      const long long ctas = (long long)num_m_tiles * s2;
      const double eff = (double)ctas / (double)(((ctas + num_sms - 1) / num_sms) * num_sms);
      if (eff > best) best = eff;
    }
    for (int s2 = 1; s2 <= smax; ++s2) {
      const long long ctas = (long long)num_m_tiles * s2;
      const double eff = (double)ctas / (double)(((ctas + num_sms - 1) / num_sms) * num_sms);
      if (eff >= best - 0.03) { splits = s2; break; }
    }
    if (const char* e = getenv("FR_GALLERY_SPLITS")) splits = std::max(1, std::min(atoi(e), n_tiles));   // A/B switch
  }
  const int tps = n_tiles > 0 ? ceil_div(n_tiles, splits) : 0;
  if (n_tiles > 0) splits = ceil_div(n_tiles, tps);
  const size_t qelems = (size_t)nq_pad * DIM;
  if (!g->q_bf16.reserve(qelems * 2) || !g->part_s.reserve((size_t)splits * nq_pad * TOPK * 4) ||
      !g->part_i.reserve((size_t)splits * nq_pad * TOPK * 4))
    return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
  const float* d_q = queries;
  if (memspace != FR_MEM_DEVICE) {
    if (!g->q_f32.reserve((size_t)nq * DIM * 4)) return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    FR_CUDA_OK(ctx, cudaMemcpyAsync(g->q_f32.p, queries, (size_t)nq * DIM * 4, cudaMemcpyHostToDevice, ctx->stream));
    d_q = g->q_f32.as<float>();
  }
  ctx->stage_begin(FR_STAGE_GALLERY);
  if (nq_pad > nq)
    FR_CUDA_OK(ctx, cudaMemsetAsync(g->q_bf16.as<bf16>() + (size_t)nq * DIM, 0, (size_t)(nq_pad - nq) * DIM * 2, ctx->stream));
  rows_to_bf16_kernel<<<148 * 2, 256, 0, ctx->stream>>>(d_q, g->q_bf16.as<bf16>(), (size_t)nq * DIM);
  ctx->launches++;
  if (fp8) {
    if (!g->q_e4m3.reserve(qelems)) return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    if (nq_pad > nq) FR_CUDA_OK(ctx, cudaMemsetAsync(g->q_e4m3.as<uint8_t>() + (size_t)nq * DIM, 0, (size_t)(nq_pad - nq) * DIM, ctx->stream));
    rows_to_e4m3_kernel<<<148 * 2, 256, 0, ctx->stream>>>(d_q, nullptr, g->q_e4m3.as<uint8_t>(), (size_t)nq * DIM);
    ctx->launches++;
  }
  if (n_tiles > 0) {
    CUtensorMap tmQ, tmG;
    const uint64_t g_rows = (uint64_t)((g->cap + GN - 1) / GN * GN);
    const bool maps_ok = fp8 ? (tc_make_map_2d_u8(&tmQ, g->q_e4m3.p, nq_pad, DIM, DIM, tc::BM) &&
                                tc_make_map_2d_u8(&tmG, g->rows8, g_rows, DIM, DIM, GN))
                             : (tc_make_map_2d(&tmQ, g->q_bf16.p, nq_pad, DIM, DIM, tc::BM) &&
                                tc_make_map_2d(&tmG, g->rows, g_rows, DIM, DIM, GN));
    if (!maps_ok)
      return fr_fail(ctx, FR_ERR_CUDA, "gallery tensor map encode failed");
    GParams p;
    p.n_rows = (int)g->size;
    p.nq = nq;
    p.num_m_tiles = num_m_tiles;
    p.n_tiles = n_tiles;
    p.tiles_per_split = tps;
    p.nq_pad = nq_pad;
    p.out_s = g->part_s.as<float>();
    p.out_i = g->part_i.as<int>();
    p.err_flag = g->err_flag;
    if (fp8) gallery_topk_kernel<true><<<num_m_tiles * splits, tc::NUM_THREADS, G_SMEM8, ctx->stream>>>(tmQ, tmG, p);
    else gallery_topk_kernel<false><<<num_m_tiles * splits, tc::NUM_THREADS, G_SMEM, ctx->stream>>>(tmQ, tmG, p);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
  }
  if (fp8) {
    // candidate set = union of at most RERANK_MAX_PARTS top-16 lists (split lists, group-merged when a small
    // query batch was spread over more splits), re-scored exactly from the bf16 rows
    int parts = n_tiles > 0 ? splits : 0;
    const float* cs = g->part_s.as<float>();
    const int* ci = g->part_i.as<int>();
    if (parts > RERANK_MAX_PARTS) {
      const int ppg = ceil_div(parts, RERANK_MAX_PARTS), groups = ceil_div(parts, ppg);
      if (!g->grp_s.reserve((size_t)groups * nq_pad * TOPK * 4) || !g->grp_i.reserve((size_t)groups * nq_pad * TOPK * 4))
        return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
      coarse_group_merge_kernel<<<dim3(ceil_div(nq, 128), groups), 128, 0, ctx->stream>>>(
          cs, ci, parts, ppg, nq, nq_pad, g->grp_s.as<float>(), g->grp_i.as<int>());
      ctx->launches++;
      ci = g->grp_i.as<int>();
      parts = groups;
    }
    rerank_kernel<<<ceil_div(nq, 8), 256, 0, ctx->stream>>>(g->q_bf16.as<bf16>(), g->rows, ci, parts, nq, nq_pad, k,
                                                            (long long)g->base, d_os, d_oi, d_rec);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
    ctx->stage_end();
    return FR_OK;
  }
  topk_merge_kernel<<<ceil_div(nq, 128), 128, 0, ctx->stream>>>(
      g->part_s.as<float>(), g->part_i.as<int>(), nullptr, n_tiles > 0 ? splits : 0, nq, nq_pad, TOPK, k,
      (long long)g->base, d_os, d_oi, nullptr, d_rec);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  ctx->stage_end();
  return FR_OK;
}

// ncclAllGather, resolved at run time: the library does not link NCCL.  A C++ host that links
// libnccl (or a process that already loaded it globally) is found through RTLD_DEFAULT, so the
// communicator the caller created and the collective we call belong to the same NCCL instance;
// otherwise libnccl.so.2 is opened.
#include <dlfcn.h>
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
static NcclAllGatherFn resolve_nccl_allgather() {
  static NcclAllGatherFn fn = [] {
    void* s = dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!s) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) s = dlsym(h, "ncclAllGather");
    }
    return reinterpret_cast<NcclAllGatherFn>(s);
  }();
  return fn;
}

extern "C" {

int fr_gallery_search(fr_gallery* g, const float* queries, int nq, int k, int memspace, float* out_scores,
                      int64_t* out_idx) {
  if (!g || !queries || !out_scores || !out_idx || nq <= 0 || k <= 0 || k > TOPK) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  float* d_os = out_scores;
  long long* d_oi = reinterpret_cast<long long*>(out_idx);
  if (memspace != FR_MEM_DEVICE) {
    if (!g->out_s.reserve((size_t)nq * k * 4) || !g->out_i.reserve((size_t)nq * k * 8))
      return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    d_os = g->out_s.as<float>();
    d_oi = g->out_i.as<long long>();
  }
  FR_CHECK(gallery_search_local(g, queries, nq, k, memspace, d_os, d_oi, nullptr));
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_idx, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_gallery_search_fp8(fr_gallery* g, const float* queries, int nq, int k, int memspace, float* out_scores,
                          int64_t* out_idx) {
  if (!g || !queries || !out_scores || !out_idx || nq <= 0 || k <= 0 || k > TOPK) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  float* d_os = out_scores;
  long long* d_oi = reinterpret_cast<long long*>(out_idx);
  if (memspace != FR_MEM_DEVICE) {
    if (!g->out_s.reserve((size_t)nq * k * 4) || !g->out_i.reserve((size_t)nq * k * 8))
      return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    d_os = g->out_s.as<float>();
    d_oi = g->out_i.as<long long>();
  }
  FR_CHECK(gallery_search_local(g, queries, nq, k, memspace, d_os, d_oi, nullptr, true));
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_idx, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_gallery_search_packed_fp8(fr_gallery* g, const float* queries, int nq, int k, int memspace, uint64_t* out_records) {
  if (!g || !queries || !out_records || nq <= 0 || k <= 0 || k > TOPK) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  if (g->base + g->size > 0xfffffffeLL) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "packed records hold 32-bit row indices");
  unsigned long long* d_rec = reinterpret_cast<unsigned long long*>(out_records);
  if (memspace != FR_MEM_DEVICE) {
    if (!g->rec_local.reserve((size_t)nq * k * 8)) return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    d_rec = g->rec_local.as<unsigned long long>();
  }
  FR_CHECK(gallery_search_local(g, queries, nq, k, memspace, nullptr, nullptr, d_rec, true));
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_records, d_rec, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_gallery_search_packed(fr_gallery* g, const float* queries, int nq, int k, int memspace, uint64_t* out_records) {
  if (!g || !queries || !out_records || nq <= 0 || k <= 0 || k > TOPK) return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  if (g->base + g->size > 0xfffffffeLL) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "packed records hold 32-bit row indices");
  unsigned long long* d_rec = reinterpret_cast<unsigned long long*>(out_records);
  if (memspace != FR_MEM_DEVICE) {
    if (!g->rec_local.reserve((size_t)nq * k * 8)) return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    d_rec = g->rec_local.as<unsigned long long>();
  }
  FR_CHECK(gallery_search_local(g, queries, nq, k, memspace, nullptr, nullptr, d_rec));
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_records, d_rec, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_topk_merge_packed(fr_ctx* ctx, const uint64_t* records, int parts, int nq, int k, int memspace,
                         float* out_scores, int64_t* out_idx) {
  if (!ctx || !records || !out_scores || !out_idx || parts <= 0 || nq <= 0 || k <= 0 || k > TOPK)
    return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad merge arguments");
  GGuard gg(ctx);
  const size_t n_in = (size_t)parts * nq * k;
  const unsigned long long* d_r = reinterpret_cast<const unsigned long long*>(records);
  float* d_os = out_scores;
  long long* d_oi = reinterpret_cast<long long*>(out_idx);
  if (memspace != FR_MEM_DEVICE) {
    if (!ctx->misc[11].reserve(n_in * 8) || !ctx->misc[8].reserve((size_t)nq * k * 4) ||
        !ctx->misc[9].reserve((size_t)nq * k * 8))
      return fr_fail(ctx, FR_ERR_CUDA, "merge allocation failed");
    FR_CUDA_OK(ctx, cudaMemcpyAsync(ctx->misc[11].p, records, n_in * 8, cudaMemcpyHostToDevice, ctx->stream));
    d_r = ctx->misc[11].as<unsigned long long>();
    d_os = ctx->misc[8].as<float>();
    d_oi = ctx->misc[9].as<long long>();
  }
  topk_merge_kernel<<<ceil_div(nq, 128), 128, 0, ctx->stream>>>(nullptr, nullptr, nullptr, parts, nq, nq, k, k, 0, d_os,
                                                               d_oi, d_r, nullptr);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_idx, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_gallery_search_sharded(fr_gallery* g, void* nccl_comm, int world, const float* queries, int nq, int k,
                              int memspace, float* out_scores, int64_t* out_idx) {
  if (!g || !nccl_comm || world <= 0 || !queries || !out_scores || !out_idx || nq <= 0 || k <= 0 || k > TOPK)
    return FR_ERR_INVALID_ARG;
  fr_ctx* ctx = g->ctx;
  GGuard gg(ctx);
  if (g->base + g->size > 0xfffffffeLL) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "packed records hold 32-bit row indices");
  NcclAllGatherFn all_gather = resolve_nccl_allgather();
  if (!all_gather) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "ncclAllGather not found (libnccl.so.2 is not loadable)");
  const size_t rec_bytes = (size_t)nq * k * 8;
  if (!g->rec_local.reserve(rec_bytes) || !g->rec_all.reserve(rec_bytes * world))
    return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
  float* d_os = out_scores;
  long long* d_oi = reinterpret_cast<long long*>(out_idx);
  if (memspace != FR_MEM_DEVICE) {
    if (!g->out_s.reserve((size_t)nq * k * 4) || !g->out_i.reserve((size_t)nq * k * 8))
      return fr_fail(ctx, FR_ERR_CUDA, "search allocation failed");
    d_os = g->out_s.as<float>();
    d_oi = g->out_i.as<long long>();
  }
  FR_CHECK(gallery_search_local(g, queries, nq, k, memspace, nullptr, nullptr, g->rec_local.as<unsigned long long>()));
  // ONE all-gather of 8-byte {score, global index} records over NVLink / NVSwitch (rank-major result)
  const int nccl_status = all_gather(g->rec_local.p, g->rec_all.p, rec_bytes, /*ncclUint8*/ 1, nccl_comm, ctx->stream);
  if (nccl_status != 0) return fr_fail(ctx, FR_ERR_CUDA, "ncclAllGather failed with status " + std::to_string(nccl_status));
  topk_merge_kernel<<<ceil_div(nq, 128), 128, 0, ctx->stream>>>(nullptr, nullptr, nullptr, world, nq, nq, k, k, 0, d_os,
                                                               d_oi, g->rec_all.as<unsigned long long>(), nullptr);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_idx, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

int fr_topk_merge(fr_ctx* ctx, const float* scores, const int64_t* idx, int parts, int nq, int k, int memspace,
                  float* out_scores, int64_t* out_idx) {
  if (!ctx || !scores || !idx || !out_scores || !out_idx || parts <= 0 || nq <= 0 || k <= 0 || k > TOPK)
    return fr_fail(ctx, FR_ERR_INVALID_ARG, "bad merge arguments");
  GGuard gg(ctx);
  const size_t n_in = (size_t)parts * nq * k;
  const float* d_s = scores;
  const long long* d_i = reinterpret_cast<const long long*>(idx);
  float* d_os = out_scores;
  long long* d_oi = reinterpret_cast<long long*>(out_idx);
  if (memspace != FR_MEM_DEVICE) {
    if (!ctx->misc[10].reserve(n_in * 4) || !ctx->misc[11].reserve(n_in * 8) ||
        !ctx->misc[8].reserve((size_t)nq * k * 4) || !ctx->misc[9].reserve((size_t)nq * k * 8))
      return fr_fail(ctx, FR_ERR_CUDA, "merge allocation failed");
    FR_CUDA_OK(ctx, cudaMemcpyAsync(ctx->misc[10].p, scores, n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(ctx->misc[11].p, idx, n_in * 8, cudaMemcpyHostToDevice, ctx->stream));
    d_s = ctx->misc[10].as<float>();
    d_i = ctx->misc[11].as<long long>();
    d_os = ctx->misc[8].as<float>();
    d_oi = ctx->misc[9].as<long long>();
  }
  topk_merge_kernel<<<ceil_div(nq, 128), 128, 0, ctx->stream>>>(d_s, nullptr, d_i, parts, nq, nq, k, k, 0, d_os, d_oi);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  if (memspace != FR_MEM_DEVICE) {
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaMemcpyAsync(out_idx, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FR_OK;
}

}  // extern "C"
