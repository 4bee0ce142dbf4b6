// tcgen05 / TMEM / TMA shift-GEMM kernel for sm_100a: the implicit-GEMM engine behind the
// IResNet-50 convolutions (K6), the FC layer (K7) and the gallery search (K9).
//
//   D[m, n] = sum over taps t, k:  A_t[m + shift_t, col_t + k] * B_t[row_t + n, k]
//
// Activations are NHWC bf16 with one shared zero halo column per row and one shared zero
// halo row per image ("padded flat layout": pixel (n,h,w) lives at flat row
// (n*Hp + h)*Wp + w, Hp = H+1, Wp = W+1).  Every out-of-image neighbour of a 3x3 stencil is
// then either a halo cell (zero) or an out-of-range row (zero-filled by TMA), so a 3x3
// convolution is nine row-shifted GEMMs over the same 2-D matrix [rows, C]; a stride-2
// convolution reads a space-to-depth copy of its input (four phase blocks of C channels per
// 2x2 cell) so that its nine taps are again plain row shifts, and the 1x1 stride-2 shortcut
// conv is one more tap accumulating into the same TMEM tile.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over output tiles):
//   warp 0   : TMA producer (one elected lane), 128x64 A tile + BNx64 B tile per stage,
//              SWIZZLE_128B, mbarrier expect_tx
//   warp 1   : MMA issuer (one elected lane): tcgen05.mma cta_group::1 kind::f16,
//              M=128, N=BN, K=16, fp32 accumulators in TMEM (double buffered: 2*BN columns)
//   warps 2-5: epilogue: tcgen05.ld 32x32b -> bias (per border class) / PReLU / residual ->
//              bf16 NHWC stores (standard, space-to-depth, + even-pixel copy) or fp32 rows
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int MAX_TAPS = 10;
constexpr int NUM_THREADS = 192;        // gallery kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int EPI_WARPS = 8;            // conv kernels: two warps per TMEM lane quarter, each half the columns
constexpr int CONV_THREADS = (2 + EPI_WARPS) * 32;

struct Tap {
  int a_src;        // 0/1: which A tensor map
  int b_src;        // 0/1: which B tensor map
  int a_row_shift;  // added to the tile's first row
  int a_col;        // first K column in A
  int b_row;        // first row in B (before adding the N-tile offset)
  int nkb;          // number of 64-wide K blocks
};

enum OutMode { OUT_STD = 0, OUT_S2D = 1, OUT_F32 = 2, OUT_TOPK = 3 };

struct Params {
  Tap taps[MAX_TAPS];
  int num_taps;
  int m_rows;       // valid output rows
  int num_m_tiles;
  int n_tiles_n;    // Cout / BN
  int H, W, Hp, Wp; // output geometry (valid extent and padded pitch)
  int Hp2, Wp2;     // geometry of the s2d / even-copy targets
  int cout;
  int bias_classes; // 1, or 9 = (top,mid,bottom) x (left,mid,right)
  int out_mode;
  const float* bias;
  const float* prelu;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  __nv_bfloat16* out_even;
  float* out_f32;
  int* err_flag;
  // split-K (FC): tile index -> (m, n, split); each split covers k_split_len K blocks and
  // writes its partial sums to out_f32 + split * split_stride (bias only in split 0)
  int k_splits;
  int k_split_len;
  long long split_stride;
  // halo mode (halo_gemm_kernel): one A block of a_rows rows per K block serves all 9 taps
  int a_rows;        // multiple of 8 (and of 16 when loaded as two boxes)
  int a_boxes;       // 1 or 2 TMA boxes per A block
  int base_off_mode; // 0: descriptor base_offset = 0; 1: base_offset = (addr >> 7) & 7
  int a_stages;      // halo mode: A blocks in flight
  // halo mode, MT == 1, streamed weights: the last, partial wave of work items is split along N so that it
  // takes a fraction of a tile time (900 tiles on 148 CTAs = 6.08 waves used to cost 7).  Items
  // [0, tail_first) are whole tiles; from tail_first on every tile is tail_split items of BN / tail_split
  // columns (tail_split in {1, 2, 4}, BN / tail_split >= 64).
  int tail_first;
  int tail_split;
  // exact division by multiply-shift: m / (Hp*Wp) = m * ceil(2^40 / (Hp*Wp)) >> 40 for m * Hp*Wp < 2^40,
  // r / Wp = umulhi(r, ceil(2^32 / Wp)) for r < Hp*Wp
  unsigned long long per_img_magic;
  uint32_t wp_magic;
  // PReLU slopes in the kernel-parameter (constant) bank: every lane of the epilogue reads the same 32
  // slopes per chunk, so they come through the constant cache instead of 8 L1 loads per thread and chunk
  int prelu_in_params;
  // halo mode, OUT_STD, one N tile: the epilogue stages the bf16 tile in shared memory (SWIZZLE_128B) and the
  // tile leaves through cp.async.bulk.tensor stores (whole 128-byte lines from the async proxy) instead of
  // lane-per-row 16-byte st.global, which cap at ~1.5 TB/s (the 64-channel layers were bound by exactly that)
  int tma_store;
  __align__(16) float prelu_c[512];
};

template <int BN> struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;  // barriers + align slack
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Blocking wait with a suspend-time hint (the warp sleeps in hardware instead of spinning and
// stealing issue slots from the working warps; ncu showed a third of all executed instructions
// in the un-hinted spin).  Bounded: a protocol bug traps (launch failure) instead of hanging.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  if (ok) return;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) {
      if (err_flag) atomicExch(err_flag, 1);
      __trap();
    }
  }
}
// true in exactly one (the lowest active) lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Pure spin (no suspend hint): for waits on the critical hand-off path, where the wake-up latency
// of a suspended warp would be paid once per pipeline stage.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity, int* err_flag) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 0xfffu) == 0xfffu) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) {
        if (err_flag) atomicExch(err_flag, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// TMA store of one box shared -> global (bulk async-group completion); rows outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(const void* p) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(p) >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 MMA with a compile-time accumulate flag (folds to UPT / !UPT: no predicate set-up in the
// single-lane issue stream)
template <bool ACC>
__device__ __forceinline__ void mma_bf16_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic) { return __umulhi(x, magic); }

template <bool WAIT = true>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  if (WAIT) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Per-row epilogue geometry (one thread = one accumulator row).  Depends only on the tile
// coordinates, not on the accumulator, so it -- and the first residual loads -- are issued BEFORE the
// epilogue warp waits for the tile's MMAs: on the 64 / 128-channel conv2 layers the residual comes from
// DRAM (~1 us) while a tile's MMAs take 0.9 us, and with only two TMEM buffers that latency was exposed.
struct EpiRow {
  bool valid, write_even, has_res;
  int cls;
  size_t out_off, even_off;
  const __nv_bfloat16* res;   // residual row (this thread's row, column n0), or null
};

__device__ __forceinline__ EpiRow epi_row(const Params& p, int m, int n0, int split) {
  EpiRow r;
  bool valid = m < p.m_rows;
  int img = 0, hp = 0, wp = 0;
  if (valid && p.out_mode != OUT_F32) {
    const int per_img = p.Hp * p.Wp;
    img = (int)(((unsigned long long)(uint32_t)m * p.per_img_magic) >> 40);
    const int rem = m - img * per_img;
    hp = (int)fast_div((uint32_t)rem, p.wp_magic);
    wp = rem - hp * p.Wp;
    valid = hp < p.H && wp < p.W;
  }
  r.cls = 0;
  if (p.bias_classes == 9)
    r.cls = (hp == 0 ? 0 : (hp == p.H - 1 ? 2 : 1)) * 3 + (wp == 0 ? 0 : (wp == p.W - 1 ? 2 : 1));
  r.out_off = 0;
  r.even_off = 0;
  r.write_even = false;
  if (p.out_mode == OUT_STD) {
    r.out_off = (size_t)m * p.cout + n0;
    if (p.out_even && valid && !(hp & 1) && !(wp & 1)) {
      r.write_even = true;
      r.even_off = ((size_t)(img * p.Hp2 + (hp >> 1)) * p.Wp2 + (wp >> 1)) * p.cout + n0;
    }
  } else if (p.out_mode == OUT_S2D) {
    r.out_off = (((size_t)(img * p.Hp2 + (hp >> 1)) * p.Wp2 + (wp >> 1)) * 4 +
                 (size_t)((hp & 1) * 2 + (wp & 1))) * p.cout + n0;
  } else {
    r.out_off = (size_t)split * p.split_stride + (size_t)m * p.cout + n0;
  }
  r.valid = valid;
  r.has_res = valid && p.residual != nullptr && p.out_mode != OUT_F32;
  r.res = r.has_res ? p.residual + (size_t)m * p.cout + n0 : nullptr;
  return r;
}

// the 32 residual values (64 bytes) of chunk c of this thread's row; streaming: keep L1 for the bias rows
__device__ __forceinline__ void epi_load_res(const EpiRow& r, int c, uint4 (&rv)[4]) {
  if (r.has_res) {
    const uint4* g = reinterpret_cast<const uint4*>(r.res + c * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) rv[i] = __ldcs(g + i);
  }
}

// Epilogue of one output tile for one thread (= one accumulator row): TMEM -> registers ->
// bias / PReLU / residual -> global.  All 32 lanes of the warp must call it (tcgen05.ld).
// rv holds the residual of chunk c_begin (epi_load_res, issued by the caller before its wait).
// stage != 0: shared address of this 128-row sub-tile's staging buffer ([BN/64 boxes][128 rows][128 B],
// SWIZZLE_128B); the OUT_STD tile goes there (zeros for halo / out-of-range rows) instead of to global memory.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const Params& p, const EpiRow& er, uint32_t taddr0, int n0, int split,
                                              int c_begin, int c_end, uint4 (&rv)[4], uint32_t stage = 0,
                                              int tile_row = 0) {
  const bool valid = er.valid;
  const float* bias = p.bias + (size_t)er.cls * p.cout + n0;
  const float bias_on = split == 0 ? 1.f : 0.f;
  const size_t out_off = er.out_off, even_off = er.even_off;
  const bool write_even = er.write_even;
  const bool has_res = er.has_res;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    // TMEM load first, then every global load the chunk needs (bias, PReLU slopes, residual:
    // none depends on the accumulator), then one wait: the latencies overlap
    uint32_t v[32];
    tmem_ld32<false>(taddr0 + c * 32, v);
    float4 b4[8], s4[8];
    uint4 rn[4];
    if (c + 1 < c_end) epi_load_res(er, c + 1, rn);   // next chunk's residual under this chunk's work
    if (valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + i * 4));
      if (p.prelu) {
        if (p.prelu_in_params) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s4[i] = *reinterpret_cast<const float4*>(&p.prelu_c[n0 + c * 32 + i * 4]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) s4[i] = __ldg(reinterpret_cast<const float4*>(p.prelu + n0 + c * 32 + i * 4));
        }
      }
    }
    tmem_wait_ld();
    if (valid) {
      float f[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[4 * i] = __uint_as_float(v[4 * i]) + b4[i].x * bias_on;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4[i].y * bias_on;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4[i].z * bias_on;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4[i].w * bias_on;
      }
      if (p.prelu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[4 * i] = f[4 * i] > 0.f ? f[4 * i] : f[4 * i] * s4[i].x;
          f[4 * i + 1] = f[4 * i + 1] > 0.f ? f[4 * i + 1] : f[4 * i + 1] * s4[i].y;
          f[4 * i + 2] = f[4 * i + 2] > 0.f ? f[4 * i + 2] : f[4 * i + 2] * s4[i].z;
          f[4 * i + 3] = f[4 * i + 3] > 0.f ? f[4 * i + 3] : f[4 * i + 3] * s4[i].w;
        }
      }
      if (p.out_mode == OUT_F32) {
        float4* o = reinterpret_cast<float4*>(p.out_f32 + out_off + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      } else {
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {rv[i].x, rv[i].y, rv[i].z, rv[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[i * 8 + 2 * j] += __uint_as_float(w[j] << 16);
              f[i * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
          }
        }
        uint4 pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pk[i].x = pack_bf16(f[i * 8 + 0], f[i * 8 + 1]);
          pk[i].y = pack_bf16(f[i * 8 + 2], f[i * 8 + 3]);
          pk[i].z = pack_bf16(f[i * 8 + 4], f[i * 8 + 5]);
          pk[i].w = pack_bf16(f[i * 8 + 6], f[i * 8 + 7]);
        }
        if (stage) {
          // 32 columns = 64 bytes = chunks [4 * (c & 1), +4) of the 128-byte row of box c / 2
          const uint32_t base = stage + (uint32_t)(c >> 1) * (128u * 128u) + (uint32_t)tile_row * 128u;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts_u4(base + (uint32_t)((((c & 1) * 4 + i) ^ (tile_row & 7)) << 4), pk[i]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(p.out + out_off + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) __stcs(o + i, pk[i]);
        }
        if (write_even) {
          uint4* oe = reinterpret_cast<uint4*>(p.out_even + even_off + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) __stcs(oe + i, pk[i]);
        }
      }
    }
    if (stage && !valid) {
      const uint32_t base = stage + (uint32_t)(c >> 1) * (128u * 128u) + (uint32_t)tile_row * 128u;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int i = 0; i < 4; ++i) sts_u4(base + (uint32_t)((((c & 1) * 4 + i) ^ (tile_row & 7)) << 4), z);
    }
    if (c + 1 < c_end && has_res) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rv[i] = rn[i];
    }
  }
}

// tile index -> (m tile, n tile, K split)
__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& m_tile, int& n_tile, int& split) {
  const int ks = p.k_splits > 1 ? p.k_splits : 1;
  split = tile % ks;
  const int t2 = tile / ks;
  n_tile = t2 % p.n_tiles_n;
  m_tile = t2 / p.n_tiles_n;
}

// ------------------------------------------------------------------------ kernel
template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1)
shift_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmB1);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.n_tiles_n * (p.k_splits > 1 ? p.k_splits : 1);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_tile, n_tile, split;
        tile_coords(p, tile, m_tile, n_tile, split);
        const int m0 = m_tile * BM, n0 = n_tile * BN;
        for (int t = 0; t < p.num_taps; ++t) {
          const Tap tp = p.taps[t];
          const CUtensorMap* ma = tp.a_src ? &tmA1 : &tmA0;
          const CUtensorMap* mb = tp.b_src ? &tmB1 : &tmB0;
          const int kb0 = p.k_splits > 1 ? split * p.k_split_len : 0;
          const int nkb = p.k_splits > 1 ? p.k_split_len : tp.nkb;
          for (int kb = kb0; kb < kb0 + nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1u, p.err_flag);
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            tma_load_2d(sa, ma, &full[stage], tp.a_col + kb * BK, m0 + tp.a_row_shift);
            tma_load_2d(sa + A_TILE_BYTES, mb, &full[stage], kb * BK, tp.b_row + n0);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the loop in uniform control flow (waits, counters and descriptor
    // arithmetic on the uniform datapath); one elected lane issues the MMAs and the commits.
    // A shared-memory descriptor carries (address >> 4) in its low 14 bits under constant upper
    // bits, so stepping stages / K steps is a 32-bit add on the low word.
    {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      const uint64_t d0 = make_smem_desc(smem);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t lo0 = (uint32_t)d0;
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int t = 0; t < p.num_taps; ++t) {
          const int nkb = p.k_splits > 1 ? p.k_split_len : p.taps[t].nkb;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full[stage], phase, p.err_flag);
            tc_fence_after();
            const uint32_t alo = lo0 + (uint32_t)stage * (uint32_t)(C::STAGE_BYTES >> 4);
            const uint32_t blo = alo + (uint32_t)(A_TILE_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
                mma_bf16(d_tmem, desc(alo + k * 2), desc(blo + k * 2), idesc, (accumulate | (uint32_t)k) ? 1u : 0u);
              tc_commit(&empty[stage]);
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (elect_one()) tc_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    constexpr int CPS = BN / 32 / (EPI_WARPS / 4);   // 32-column chunks per warp
    const int c_begin = ((warp - 2) >> 2) * CPS;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int m_tile, n_tile, split;
      tile_coords(p, tile, m_tile, n_tile, split);
      const uint32_t acc = it & 1u;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const EpiRow er = epi_row(p, m_tile * BM + row, n_tile * BN, split);
      uint4 rv[4];
      epi_load_res(er, c_begin, rv);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      epilogue_tile<BN>(p, er, taddr0, n_tile * BN, split, c_begin, c_begin + CPS, rv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- halo-mode kernel
// 3x3 stride-1 convolutions only.  Instead of re-loading the 128-row A tile once per tap
// (9 x 16 KB per K block), ONE block of a_rows = 128 + 2*Wp + 2 (rounded up to 8) rows is
// loaded per K block and the nine taps address it through row-shifted shared-memory
// descriptors (start address + (r*Wp + s) * 128 B; SWIZZLE_128B is a function of the
// shared-memory address bits, which TMA used when it wrote the block; verified on B200: the
// descriptor base_offset field must stay 0).  Weights stream
// through their own ring.  Cuts the L2->SMEM traffic of the feed-bound layers 1.4-2.1x.
template <int BN, int MT, bool RESB> struct HaloCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int MAX_A_STAGES = 6;   // runtime p.a_stages: as many A blocks in flight as shared memory allows
  static constexpr int B_STAGES = 4;
  static constexpr int ACC_STAGES = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = ACC_STAGES * MT * BN < 32 ? 32 : ACC_STAGES * MT * BN;
  static constexpr int B_BYTES = RESB ? 9 * B_TILE_BYTES : B_STAGES * B_TILE_BYTES;
  // TMA-store staging: two buffers (the store of sub-tile i drains while sub-tile i + 1 is written) of BN/64 boxes
  static constexpr int STAGE_OUT_BYTES = (BN / 64) * 128 * 128;
  static int smem_bytes(int a_rows, int a_stages, bool tma_store = false) {
    return a_stages * a_rows * 128 + B_BYTES + 256 + 1024 + (tma_store ? 2 * STAGE_OUT_BYTES + 1024 : 0);
  }
  static int pick_a_stages(int a_rows, bool tma_store = false) {
    int s = 2;
    while (s < MAX_A_STAGES && smem_bytes(a_rows, s + 1, tma_store) <= 227 * 1024) ++s;
    return s;
  }
};

// MT   : M tiles (of 128 rows) per CTA iteration; they share one A block of
//        MT*128 + 2*Wp + 2 rows and every streamed weight tile (halves the weight traffic)
// RESB : Cin == 64 only: all nine weight tiles stay resident in shared memory
template <int BN, int MT, bool RESB>
__global__ void __launch_bounds__(CONV_THREADS, 1)
halo_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBs, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ Params p) {
  using C = HaloCfg<BN, MT, RESB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_bytes = p.a_rows * 128;              // multiple of 1024
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.a_stages * a_bytes;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sB + C::B_BYTES);
  uint64_t* aempty = afull + C::MAX_A_STAGES;
  uint64_t* bfull = aempty + C::MAX_A_STAGES;
  uint64_t* bempty = bfull + C::B_STAGES;
  uint64_t* tfull = bempty + C::B_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resfull + 1);
  // TMA-store staging (1024-byte aligned for SWIZZLE_128B), after the 256-byte barrier block
  uint8_t* sOut = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sB + C::B_BYTES + 256) + 1023) &
                                             ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    if (p.tma_store) prefetch_tmap(&tmOut);
    for (int s = 0; s < C::MAX_A_STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    mbar_init(resfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmBs);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_super = (p.num_m_tiles + MT - 1) / MT;
  const int tail_split = (MT == 1 && !RESB && p.tail_split > 1) ? p.tail_split : 1;
  const int tail_first = tail_split > 1 ? p.tail_first : num_super * p.n_tiles_n;
  const int total_tiles = tail_first + (num_super * p.n_tiles_n - tail_first) * tail_split;   // work items
  const int nkb = p.taps[0].nkb;
  const int wp = p.Wp;
  const int box_rows = p.a_rows / p.a_boxes;
  // work item -> first output row, first output column, number of columns
  auto item_coords = [&](int item, int& m0, int& n0, int& nlen) {
    int tile = item, sub = 0;
    nlen = BN;
    if (item >= tail_first) {
      const int j = item - tail_first;
      tile = tail_first + j / tail_split;
      sub = j % tail_split;
      nlen = BN / tail_split;
    }
    m0 = (tile / p.n_tiles_n) * (BM * MT);
    n0 = (tile % p.n_tiles_n) * BN + sub * nlen;
  };

  if (warp == 0) {
    if (lane == 0) {
      if (RESB) {   // n_tiles_n == 1 and nkb == 1: the whole weight tensor is 9 tiles
        mbar_expect_tx(resfull, 9 * C::B_TILE_BYTES);
        for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * C::B_TILE_BYTES, &tmB, resfull, 0, t * p.cout);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m0, n0, nlen;
        item_coords(tile, m0, n0, nlen);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&aempty[sa], pa ^ 1u, p.err_flag);
          mbar_expect_tx(&afull[sa], (uint32_t)a_bytes);
          for (int bx = 0; bx < p.a_boxes; ++bx)
            tma_load_2d(sA + sa * a_bytes + bx * box_rows * 128, &tmA, &afull[sa], kb * BK,
                        m0 - wp - 1 + bx * box_rows);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
          if (!RESB) {
            for (int t = 0; t < 9; ++t) {
              mbar_wait(&bempty[sb], pb ^ 1u, p.err_flag);
              if (nlen == BN) {
                mbar_expect_tx(&bfull[sb], C::B_TILE_BYTES);
                tma_load_2d(sB + sb * C::B_TILE_BYTES, &tmB, &bfull[sb], kb * BK, t * p.cout + n0);
              } else {   // a split tail item: nlen / 64 boxes of 64 weight rows
                mbar_expect_tx(&bfull[sb], (uint32_t)(nlen * BK * 2));
                for (int r = 0; r < nlen; r += 64)
                  tma_load_2d(sB + sb * C::B_TILE_BYTES + r * BK * 2, &tmBs, &bfull[sb], kb * BK, t * p.cout + n0 + r);
              }
              if (++sb == C::B_STAGES) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (see shift_gemm_kernel)
    {
      constexpr uint32_t idesc_full = make_idesc(BM, BN);
      const uint32_t idesc_tail = make_idesc(BM, BN / tail_split);
      const uint64_t d0 = make_smem_desc(sA);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;
      const uint32_t b_lo0 = (uint32_t)make_smem_desc(sB);
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      const uint32_t a_step = (uint32_t)a_bytes >> 4;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, it = 0;
      if (RESB) {
        mbar_wait(resfull, 0, p.err_flag);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = C::ACC_STAGES == 2 ? (it & 1u) : 0u;
        const uint32_t acc_phase = C::ACC_STAGES == 2 ? ((it >> 1) & 1u) : (it & 1u);
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * BN);
        const uint32_t idesc = tile >= tail_first ? idesc_tail : idesc_full;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&afull[sa], pa, p.err_flag);
          tc_fence_after();
          const uint32_t ablk = a_lo0 + (uint32_t)sa * a_step;
          if (RESB) {
            // Cin == 64: one K block, resident weights, nothing to wait for between taps.  All 36 * MT MMAs of
            // the tile are issued back to back from ONE elected region: three row bases (dy = 0, 1, 2), every
            // other descriptor offset and every accumulate flag is an immediate.  (N = 64 MMAs last 32 cycles:
            // the per-tap elect / R2UR / branch sequence of the generic loop, ~32 instructions per 4 MMAs, left
            // the tensor pipe idle half of the time: 49.5 % active in ncu.)
            if (elect_one()) {
              const uint32_t arow[3] = {ablk, ablk + (uint32_t)wp * 8u, ablk + (uint32_t)wp * 16u};
#pragma unroll
              for (int t = 0; t < 9; ++t) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t adesc = arow[t / 3] + (uint32_t)((t % 3) * 8 + mt * BM * 8);
                  const uint32_t btile = b_lo0 + (uint32_t)t * (uint32_t)(C::B_TILE_BYTES >> 4);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
                    if (t == 0 && k == 0) mma_bf16_c<false>(d_tmem + mt * BN, desc(adesc + k * 2), desc(btile + k * 2), idesc_full);
                    else mma_bf16_c<true>(d_tmem + mt * BN, desc(adesc + k * 2), desc(btile + k * 2), idesc_full);
                  }
                }
              }
              tc_commit(&aempty[sa]);
            }
            __syncwarp();
            if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
            continue;
          }
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            uint32_t btile;
            if (RESB) {
              btile = b_lo0 + (uint32_t)t * (uint32_t)(C::B_TILE_BYTES >> 4);
            } else {
              mbar_wait(&bfull[sb], pb, p.err_flag);
              tc_fence_after();
              btile = b_lo0 + (uint32_t)sb * (uint32_t)(C::B_TILE_BYTES >> 4);
            }
            // a row of the A block is 128 bytes = 8 descriptor units
            const uint32_t off_rows = (uint32_t)((t / 3) * wp + (t % 3)) * 8u;
            const uint32_t accumulate = (kb | t) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t adesc = ablk + off_rows + (uint32_t)(mt * BM) * 8u;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  mma_bf16(d_tmem + mt * BN, desc(adesc + k * 2), desc(btile + k * 2), idesc,
                           (accumulate | (uint32_t)k) ? 1u : 0u);
              }
              if (!RESB) tc_commit(&bempty[sb]);
            }
            __syncwarp();
            if (!RESB) {
              if (++sb == C::B_STAGES) { sb = 0; pb ^= 1u; }
            }
          }
          if (elect_one()) tc_commit(&aempty[sa]);
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
        }
        if (elect_one()) tc_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    constexpr int CPS = BN / 32 / (EPI_WARPS / 4);   // 32-column chunks per warp
    const int c_begin = ((warp - 2) >> 2) * CPS;
    uint32_t it = 0, st_count = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int m0, n0, nlen;
      item_coords(tile, m0, n0, nlen);
      const int cps = nlen == BN ? CPS : nlen / 32 / (EPI_WARPS / 4);      // this item's chunks per warp
      const int cb = nlen == BN ? c_begin : ((warp - 2) >> 2) * cps;
      const uint32_t acc = C::ACC_STAGES == 2 ? (it & 1u) : 0u;
      const uint32_t acc_phase = C::ACC_STAGES == 2 ? ((it >> 1) & 1u) : (it & 1u);
      // row geometry + the first residual loads go out before the wait for this tile's MMAs
      EpiRow er = epi_row(p, m0 + row, n0, 0);
      uint4 rv[4];
      epi_load_res(er, cb, rv);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const bool staged = p.tma_store && nlen == BN;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t taddr0 = tmem_base + acc * (MT * BN) + mt * BN + ((uint32_t)(q * 32) << 16);
        // staged: sub-tile st uses staging buffer st & 1.  The issuing thread (warp 2, lane 0) has waited until
        // the store that last read this buffer (two sub-tiles ago) is done with it; barrier 1 tells everyone.
        if (staged) named_bar_sync(1, EPI_WARPS * 32);
        const uint32_t stage = staged ? smem_u32(sOut) + (st_count & 1u) * (uint32_t)C::STAGE_OUT_BYTES : 0u;
        epilogue_tile<BN>(p, er, taddr0, n0, 0, cb, cb + cps, rv, stage, row);
        if (staged) {
          fence_proxy_async_smem();                 // generic-proxy writes -> visible to the TMA engine
          named_bar_sync(2, EPI_WARPS * 32);
          if (warp == 2 && lane == 0) {
#pragma unroll
            for (int bx = 0; bx < BN / 64; ++bx)
              tma_store_2d(&tmOut, sOut + (st_count & 1u) * C::STAGE_OUT_BYTES + bx * (128 * 128), n0 + bx * 64,
                           m0 + mt * BM);
            tma_store_commit();
            tma_store_wait_read<1>();               // the OTHER buffer (previous sub-tile's store) is free again
          }
          ++st_count;
        }
        if (mt + 1 < MT) {
          er = epi_row(p, m0 + (mt + 1) * BM + row, n0, 0);
          epi_load_res(er, cb, rv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    // the bulk stores must have finished reading shared memory before the CTA exits
    if (p.tma_store && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}


// ------------------------------------------------------- 2-CTA halo-mode kernel
// The 256 / 512-channel 3x3 stride-1 convolutions (28 + 4 launches, half of the trunk's time) are not
// bound by the tensor pipe but by the L2 -> shared-memory feed: with cta_group::1 every CTA streams the
// whole 256 x 64 weight tile of every tap (9 x 32 KB per 20 KB of activations per K block, 67 B/clk/SM
// at the MMA's own pace; measured tensor-pipe activity 73-75 %).  A CTA PAIR (cluster of 2, one TPC)
// executes one tcgen05.mma.cta_group::2 of M = 256: each CTA holds its own 128-row activation block and
// only HALF of the weight tile (its N/2 rows); the tensor cores of the two SMs exchange the halves.  Per
// SM that is 9 x 16 KB of weights per K block (36 B/clk) and 8 KB of shared-memory operand reads per
// 128-cycle MMA instead of 12 KB.
//   * both CTAs run the same code on the same shared-memory layout; work item = 256 rows x BN columns;
//   * producers: each CTA loads its A block and its half of B with cp.async.bulk.tensor...cta_group::2,
//     completing on the LEADER's (cluster rank 0) full barriers, which the leader arms for both halves;
//   * MMA: issued by the leader only; tcgen05.commit...multicast::cluster releases the stage / publishes
//     the accumulator in BOTH CTAs (each waits on its own barrier copy);
//   * epilogue: each CTA drains its own 128 TMEM lanes; the peer's warps arrive remotely on the leader's
//     tempty barrier.
// Tail items (last partial wave) are split along N like in halo_gemm_kernel: N = BN / tail_split, each
// CTA loading N/2 weight rows through the map with the matching box height.
// Used for BN = 256 only.  Measured on B200 (profiles/r2_two_cta_mask_sweep.txt): the pair pays where the
// MMA is long (N = 256: 128 clk per instruction): trunk 6.01 -> 5.76 ms; with N = 64 MMAs (32 clk) the
// 64-channel layers got 40 % SLOWER (the pair's per-instruction hand-shake dominates), N = 128 gave nothing,
// and the stride-2 convs on pairs (per-tap A tiles, half weight tiles) measured 5.835 vs 5.836 ms.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the mbarrier of the pair's leader CTA (address with
// the CTA-rank bit cleared, like CUTLASS's SM100_TMA_2SM_LOAD)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of all prior MMAs of this thread: one arrival on the barrier at this offset in both CTAs
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(rank)
      : "memory");
}

template <int BN> struct Halo2Cfg {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;   // this CTA's half of one weight tile
  static constexpr int MAX_A_STAGES = 6;
  static constexpr int B_STAGES = 6;
  static constexpr int B_BYTES = B_STAGES * B_HALF_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // double-buffered accumulators
  static int smem_bytes(int a_rows, int a_stages) { return a_stages * a_rows * 128 + B_BYTES + 256 + 1024; }
  static int pick_a_stages(int a_rows) {
    int s = 2;
    while (s < MAX_A_STAGES && smem_bytes(a_rows, s + 1) <= 200 * 1024) ++s;
    return s;
  }
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
halo_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmB4,
                  const __grid_constant__ Params p) {
  using C = Halo2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_bytes = p.a_rows * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.a_stages * a_bytes;
  uint64_t* afull = reinterpret_cast<uint64_t*>(sB + C::B_BYTES);
  uint64_t* aempty = afull + C::MAX_A_STAGES;
  uint64_t* bfull = aempty + C::MAX_A_STAGES;
  uint64_t* bempty = bfull + C::B_STAGES;
  uint64_t* tfull = bempty + C::B_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::MAX_A_STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_super = (p.num_m_tiles + 1) / 2;
  const int tail_split = p.tail_split > 1 ? p.tail_split : 1;
  const int tail_first = tail_split > 1 ? p.tail_first : num_super * p.n_tiles_n;
  const int total_items = tail_first + (num_super * p.n_tiles_n - tail_first) * tail_split;
  const int nkb = p.taps[0].nkb;
  const int wp = p.Wp;
  const int box_rows = p.a_rows / p.a_boxes;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  auto item_coords = [&](int item, int& m0, int& n0, int& nlen) {
    int tile = item, sub = 0;
    nlen = BN;
    if (item >= tail_first) {
      const int j = item - tail_first;
      tile = tail_first + j / tail_split;
      sub = j % tail_split;
      nlen = BN / tail_split;
    }
    m0 = (tile / p.n_tiles_n) * (2 * BM) + (int)rank * BM;   // this CTA's 128 rows of the 256-row item
    n0 = (tile % p.n_tiles_n) * BN + sub * nlen;
  };

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        int m0, n0, nlen;
        item_coords(item, m0, n0, nlen);
        const int half = nlen / 2;                              // weight rows this CTA loads per tap
        const CUtensorMap* mb = nlen == BN ? &tmB : (nlen == BN / 2 ? &tmB2 : &tmB4);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&aempty[sa], pa ^ 1u, p.err_flag);
          if (leader) mbar_expect_tx(&afull[sa], 2u * (uint32_t)a_bytes);
          for (int bx = 0; bx < p.a_boxes; ++bx)
            tma_load_2d_2sm(sA + sa * a_bytes + bx * box_rows * 128, &tmA, &afull[sa], kb * BK,
                            m0 - wp - 1 + bx * box_rows);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&bempty[sb], pb ^ 1u, p.err_flag);
            if (leader) mbar_expect_tx(&bfull[sb], 2u * (uint32_t)(half * BK * 2));
            tma_load_2d_2sm(sB + sb * C::B_HALF_BYTES, mb, &bfull[sb], kb * BK, t * p.cout + n0 + (int)rank * half);
            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const uint64_t d0 = make_smem_desc(sA);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;
      const uint32_t b_lo0 = (uint32_t)make_smem_desc(sB);
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      const uint32_t a_step = (uint32_t)a_bytes >> 4;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, it = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int nlen = item >= tail_first ? BN / tail_split : BN;
        const uint32_t idesc = make_idesc(2 * BM, nlen);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&afull[sa], pa, p.err_flag);
          tc_fence_after();
          const uint32_t ablk = a_lo0 + (uint32_t)sa * a_step;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&bfull[sb], pb, p.err_flag);
            tc_fence_after();
            const uint32_t btile = b_lo0 + (uint32_t)sb * (uint32_t)(C::B_HALF_BYTES >> 4);
            const uint32_t adesc = ablk + (uint32_t)((t / 3) * wp + (t % 3)) * 8u;
            const uint32_t accumulate = (kb | t) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                mma_bf16_2sm(d_tmem, desc(adesc + k * 2), desc(btile + k * 2), idesc, (accumulate | (uint32_t)k) ? 1u : 0u);
              tc_commit_2sm(&bempty[sb]);
            }
            __syncwarp();
            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1u; }
          }
          if (elect_one()) tc_commit_2sm(&aempty[sa]);
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
        }
        if (elect_one()) tc_commit_2sm(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    constexpr int CPS = BN / 32 / (EPI_WARPS / 4);
    const int c_begin = ((warp - 2) >> 2) * CPS;
    uint32_t it = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
      int m0, n0, nlen;
      item_coords(item, m0, n0, nlen);
      const int cps = nlen == BN ? CPS : nlen / 32 / (EPI_WARPS / 4);
      const int cb = nlen == BN ? c_begin : ((warp - 2) >> 2) * cps;
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      const EpiRow er = epi_row(p, m0 + row, n0, 0);
      uint4 rv[4];
      epi_load_res(er, cb, rv);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      if (cps > 0) epilogue_tile<BN>(p, er, taddr0, n0, 0, cb, cb + cps, rv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[acc]);
        else mbar_arrive_cluster(&tempty[acc], 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer may still be signalling / reading this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}


}  // namespace tc
