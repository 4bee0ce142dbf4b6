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
// (K = channels contiguous).  An output tile is TH x TW pixels (8x16, or 6x20 on the 40^2 /
// 20^2 maps; MMA M = 128 rows) and K is consumed in blocks of KC channels (32, or 16):
//   warp 0   : TMA (3-D tiled map [C, W, N*H], OOB zero fill = the conv padding) of the input
//              halo box ((TH-1)*stride+3) x ((TW-1)*stride+3) x KC into an input ring;
//   warps 7-14 (converters): build the A operand from that box in shared memory -- the
//              depthwise 3x3 + bias + ReLU evaluated on the fly with a sliding register window
//              (the depthwise result never exists in HBM), or an im2col tap of a dense 3x3,
//              or the plain 1x1 input -- split it into tf32 hi/lo and store both as K-major
//              SWIZZLE_128B tiles; fence.proxy.async + mbarrier arrive;
//   warp 6   : TMA of the (host-pre-split) hi/lo weight tiles into the same A/B stage;
//   warp 1   : one lane issues 3 x tcgen05.mma.kind::tf32 (M=128, N=16..160, K=8) per K step;
//   warps 2-5: epilogue, tcgen05.ld -> sum of the partial accumulators -> bias / ReLU /
//              top-down upsample-add / accumulate -> fp32 NHWC float4 stores, or the
//              anchor-major score (sigmoid) / bbox / kps records.
// The tensor core truncates when it adds each K=8 product group into the fp32 accumulator
// (~2^-24 |D| bias per MMA, measured: 8e-5 at the kps head with one accumulator), so the two
// cross terms get their own accumulator and the hi*hi products alternate between two; the
// epilogue adds the partial sums with round-to-nearest (measured 2e-5 = 3.4e-4 px).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "tc_gemm.cuh"

bool tc_make_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t pitch_elems, uint32_t box_rows);   // k_iresnet.cu
bool tc_make_map_3d_f32(CUtensorMap* map, const void* base, uint64_t c, uint64_t w, uint64_t rows,
                        uint32_t box_c, uint32_t box_w, uint32_t box_rows);   // k_iresnet.cu
bool tc_make_map_2d_u64(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                        uint32_t box_rows);                                     // k_iresnet.cu

namespace {

using tc::make_smem_desc;
using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mbar_wait_spin;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tma_load_2d;

constexpr int DET = FR_DET_SIZE;
constexpr int TM = 128;                 // MMA M (tile rows; TH*TW of them are pixels)
constexpr int A_BYTES = TM * 128;       // one 128 x 32 fp32 operand tile
constexpr int MAX_STAGES = 8;
constexpr int CV_WARPS = 8;             // converter warps
constexpr int FIRST_CV_WARP = 7;        // 0 input TMA, 1 MMA, 2-5 epilogue, 6 weight TMA
constexpr int THREADS = (FIRST_CV_WARP + CV_WARPS) * 32;

enum { LD_PW = 0, LD_DW = 1, LD_IM2COL = 2 };
enum { EPI_STD = 0, EPI_HEAD = 1 };

struct SepParams {
  float* out;           // fp32 NHWC [n][hout][wout][cout]
  const float* dw_w;    // [9][cin] depthwise weights, tap major
  const float* dw_b;    // [cin]
  const float* bias;    // [npad_total]
  const float* add_up;  // optional [n][hout/2][wout/2][cout], nearest-upsampled and added
  float* score;
  float* bbox;
  float* kps;
  int cin, cout;
  int hin, hout, wout, stride;
  int rows_out;         // n * hout (image rows are merged into one axis)
  int mode, epi, relu, accumulate;
  int kc;               // channels per input box / K block (16 or 32)
  int n_in;             // input boxes (channel blocks) per tile
  int taps;             // A K-blocks built from one input box: 1 (1x1, depthwise) or 9 (dense 3x3)
  int nt;               // MMA N (per N tile)
  int n_tiles_n;
  int npad_total;       // rows of the hi half of the packed weights
  int th, tw;           // output tile = th x tw pixels; tile row = ty*tw + tx
  int tiles_x;
  uint32_t tiles_x_magic, hout_magic;   // ceil(2^32 / d) for exact small-range division (0: d == 1)
  int n_shift;          // log2(n_tiles_n)
  int num_m_tiles;
  int bw, bh, halo;     // input box (pixels) and its halo (1 for 3x3 stencils, 0 for 1x1)
  int in_bytes;         // bytes per input stage (rounded up to 1024)
  int in_merged;        // cin == kc == 16: the map is 2-D over (W * 8 u64, N*H): one request per box row
  int tail8;            // depthwise, kc = 32, stride 1, cin % 32 == 8: the last K block has 8 real channels; its box is
                        // loaded 8 channels wide (tmInTail) and converted one pixel x one chunk per thread
  int s_in, s_ab;       // ring depths
  int tmem_cols;
  int nbig;             // hi*hi accumulators per tile (k steps alternate between them)
  int acc_stages;       // 2 = TMEM double buffered, 1 = single
  long long* dbg;       // debug timeline [5 roles][256 slots][2] of clock64 (CTA 0 only), or null
  int* err_flag;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// explicit shared-space accesses on 32-bit shared addresses (the generic pointer arithmetic
// through the ring structs otherwise compiles to generic LD/ST)
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// round-to-nearest (ties away) to the 10-bit tf32 mantissa: what cvt.rna.tf32.f32 does for finite
// inputs, in two integer instructions (the PTX instruction expands to a NaN/Inf-safe sequence)
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void dbg_stamp(const SepParams& p, int role, uint32_t idx, int ev) {
  if (p.dbg && blockIdx.x == 0 && idx < 256) p.dbg[(role * 256 + idx) * 2 + ev] = clock64();
}

// x / d for d >= 1 and x * d < 2^32, with m = ceil(2^32 / d) (m == 0 encodes d == 1)
__device__ __forceinline__ int fastdiv(int x, uint32_t m) { return m ? (int)__umulhi((uint32_t)x, m) : x; }

struct TileCoord {
  int n_tile, rb, txi, gy0, y0;
};
// tile index -> N tile, row block, x tile, first merged output row and its row inside the image
__device__ __forceinline__ TileCoord tile_coord(const SepParams& p, int tile) {
  TileCoord c;
  const int m_tile = tile >> p.n_shift;
  c.n_tile = tile - (m_tile << p.n_shift);
  c.rb = fastdiv(m_tile, p.tiles_x_magic);
  c.txi = m_tile - c.rb * p.tiles_x;
  c.gy0 = c.rb * p.th;
  c.y0 = c.gy0 - fastdiv(c.gy0, p.hout_magic) * p.hout;
  return c;
}

// row-major 128-byte rows, 16-byte chunk index XOR (row & 7): what TMA / UMMA call SWIZZLE_128B
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ void split_store(uint32_t hi, uint32_t lo, uint32_t off, const float4& v) {
  float4 h, l;
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  // the residual is rounded (not left to the tensor core's truncation) to tf32 as well
  l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
  sts4(hi + off, h);
  sts4(lo + off, l);
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

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// 16 accumulator columns of one row: hi*hi partial sum(s) + the cross-term accumulator, added
// in fp32 with round-to-nearest ((big0 + big1) + small).  All TMEM loads are issued before the
// single wait.
__device__ __forceinline__ void ld_sum16(uint32_t taddr, uint32_t small_off, int nbig, uint32_t nt, uint32_t (&v)[16]) {
  uint32_t s[16], b1[16];
  tmem_ld16_nowait(taddr, v);
  tmem_ld16_nowait(taddr + small_off, s);
  if (nbig == 2) tmem_ld16_nowait(taddr + nt, b1);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (nbig == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(b1[i]));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(s[i]));
}

// five partial accumulators at column offsets 0 (hi*hi, even k), nt (hi*lo, even k), 2nt (hi*hi, odd k),
// 3nt (hi*lo, odd k), 4nt (lo*hi): (big0 + big1) + ((small0 + small1) + small2)
__device__ __forceinline__ void ld_sum16x5(uint32_t taddr, uint32_t nt, uint32_t (&v)[16]) {
  uint32_t s0[16], b1[16], s1[16], s2[16];
  tmem_ld16_nowait(taddr, v);
  tmem_ld16_nowait(taddr + nt, s0);
  tmem_ld16_nowait(taddr + 2 * nt, b1);
  tmem_ld16_nowait(taddr + 3 * nt, s1);
  tmem_ld16_nowait(taddr + 4 * nt, s2);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = __float_as_uint((__uint_as_float(v[i]) + __uint_as_float(b1[i])) +
                           ((__uint_as_float(s0[i]) + __uint_as_float(s1[i])) + __uint_as_float(s2[i])));
}

// tcgen05.mma.kind::tf32 with a compile-time accumulate flag (folds to UPT / !UPT: no predicate
// set-up in the single-lane issue stream)
template <bool ACC>
__device__ __forceinline__ void mma_tf32_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}

// One tap (KS k-steps of K = 8) of the halo 3x3 kernel: 2 MMAs per k-step.  The B tile of a stage is
// [w_hi | w_lo] (2nt contiguous rows), so x_hi * [w_hi | w_lo] is ONE MMA of N = 2nt that reads the A
// tile once for both products (the N <= 32 MMAs are bound by their shared-memory operand reads):
// columns [0, nt) of its accumulator are hi*hi, [nt, 2nt) hi*lo; k steps alternate between two such
// accumulators (offsets 0 and 2nt), x_lo * w_hi goes to a fifth region at 4nt.  Fully unrolled,
// descriptor low words stepped by immediates.  FIRST: the tile's first tap overwrites the accumulators.
template <int KS, bool FIRST>
__device__ __forceinline__ void c3_issue_tap(uint32_t d_base, uint32_t nt, uint32_t desc_hi, uint32_t ahi, uint32_t alo,
                                             uint32_t bhi, uint32_t idesc, uint32_t idesc2) {
  auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const uint32_t o = (uint32_t)(k * 2);
    if (FIRST && k == 0) {
      mma_tf32_c<false>(d_base + 4 * nt, desc(alo + o), desc(bhi + o), idesc);
      mma_tf32_c<false>(d_base, desc(ahi + o), desc(bhi + o), idesc2);
    } else {
      mma_tf32_c<true>(d_base + 4 * nt, desc(alo + o), desc(bhi + o), idesc);
      if (FIRST && k == 1) mma_tf32_c<false>(d_base + 2 * nt, desc(ahi + o), desc(bhi + o), idesc2);
      else mma_tf32_c<true>(d_base + (uint32_t)(k & 1) * 2 * nt, desc(ahi + o), desc(bhi + o), idesc2);
    }
  }
}

struct Rings {
  uint8_t* in;          // s_in input boxes
  uint8_t* ab;          // s_ab x (A_hi, A_lo, B_hi, B_lo)
  uint64_t* full_in;
  uint64_t* empty_in;
  uint64_t* full_ab;
  uint64_t* empty_ab;
  int ab_bytes;
};

// Converter warps: input box (shared memory, written by TMA) -> A operand hi/lo tiles.
//   KC = 32: a thread owns 4 consecutive output pixels of one row x one 16-byte channel chunk
//            (a quarter warp = the 8 chunks of one pixel = 128 contiguous bytes: no conflicts);
//   KC = 16: 2 consecutive pixels x one chunk; a quarter warp = 4 chunks x 2 adjacent tile rows
//            (the host makes the box width odd for stride 1 so those rows hit different banks).
template <int KC, int STRIDE, int MODE>
__device__ __forceinline__ void converter_loop(const SepParams& p, const Rings& r, const float* s_dw, int t,
                                               int lane) {
  constexpr int PXT = KC == 32 ? 4 : 2;
  constexpr int PP = KC * 4;                        // pixel pitch inside the box (bytes)
  constexpr int NV = (PXT - 1) * STRIDE + 3;        // box columns a thread's window spans
  const int xgroups = p.tw / PXT;
  int chunk, ty, xg;
  if (KC == 32) {
    chunk = t & 7;
    const int g = t >> 3;
    ty = g / xgroups;
    xg = g - ty * xgroups;
  } else {
    chunk = t & 3;
    const int g = t >> 3;
    const int typ = g / xgroups;
    xg = g - typ * xgroups;
    ty = 2 * typ + ((t >> 2) & 1);
  }
  const bool active = ty < p.th;
  const int row0 = ty * p.tw + xg * PXT;            // tile row of the first owned pixel
  uint32_t soff[PXT];
#pragma unroll
  for (int j = 0; j < PXT; ++j) soff[j] = sw128(row0 + j, chunk);
  // byte offset of the thread's window inside the box (row rr adds rr * bw * PP)
  const int box_off = (MODE == LD_PW ? (ty * p.bw + xg * PXT) : (ty * STRIDE * p.bw + xg * PXT * STRIDE)) * PP + chunk * 16;
  const int row_pitch = p.bw * PP;
  const uint32_t in_base = smem_u32(r.in), ab_base = smem_u32(r.ab);
  const int total_tiles = p.num_m_tiles << p.n_shift;
  // depthwise weights of this thread's 4 channels: [tap][cin] + bias row, staged in shared memory
  const uint32_t dw_base = smem_u32(s_dw);
  float4 w[9], b4 = zero4();
  auto load_dw = [&](int k0) {
    const bool kin = k0 < p.cin;
#pragma unroll
    for (int q = 0; q < 9; ++q) w[q] = kin ? lds4(dw_base + (uint32_t)((q * p.cin + k0) * 4)) : zero4();
    b4 = kin ? lds4(dw_base + (uint32_t)((9 * p.cin + k0) * 4)) : zero4();
  };
  if (MODE == LD_DW) load_dw(chunk * 4);
  // tail block (p.tail8): 128 pixels x 2 chunks = one (pixel, chunk) per thread, box pixel pitch 32 bytes; a
  // quarter warp reads 4 adjacent pixels = 128 contiguous bytes
  const int t_row = t >> 1, t_chunk = t & 1;
  const int t_ty = t_row / p.tw, t_tx = t_row - t_ty * p.tw;
  const bool t_active = t_ty < p.th;
  const uint32_t t_soff = sw128(t_row, t_chunk);
  const int t_box_off = (t_ty * p.bw + t_tx) * 32 + t_chunk * 16;
  int si = 0, sa = 0;
  uint32_t ph_in = 0, ph_ab = 0, dbg_kb = 0, dbg_box = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TileCoord tc = tile_coord(p, tile);
    const bool pix_ok = active && tc.gy0 + ty < p.rows_out;
    int y = tc.y0 + ty;
    if (y >= p.hout) y -= p.hout;
    for (int i = 0; i < p.n_in; ++i) {
      mbar_wait(&r.full_in[si], ph_in, p.err_flag);
      if (t == 0) dbg_stamp(p, 3, dbg_box++, 1);
      const uint32_t box = in_base + (uint32_t)(si * p.in_bytes + box_off);
      const int k0 = i * KC + chunk * 4;
      const bool kv = k0 < p.cin && pix_ok;
      if (KC == 32 && STRIDE == 1 && MODE == LD_DW && p.tail8 && i == p.n_in - 1) {
        load_dw(i * KC + t_chunk * 4);
        mbar_wait_spin(&r.empty_ab[sa], ph_ab ^ 1u, p.err_flag);
        if (t == 0) dbg_stamp(p, 0, dbg_kb, 0);
        if (t_active) {
          float4 acc = zero4();
          if (tc.gy0 + t_ty < p.rows_out) {
            int yy = tc.y0 + t_ty;
            if (yy >= p.hout) yy -= p.hout;
            acc = b4;
            const uint32_t tb = in_base + (uint32_t)(si * p.in_bytes + t_box_off);
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
              const int iy = yy - 1 + rr;
              if (iy >= 0 && iy < p.hin) {
                const uint32_t rb = tb + (uint32_t)(rr * p.bw * 32);
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) fma4(acc, lds4(rb + s2 * 32), w[rr * 3 + s2]);
              }
            }
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
          }
          const uint32_t hi = ab_base + (uint32_t)(sa * r.ab_bytes);
          split_store(hi, hi + A_BYTES, t_soff, acc);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&r.full_ab[sa]);
        if (t == 0) dbg_stamp(p, 0, dbg_kb++, 1);
        if (++sa == p.s_ab) { sa = 0; ph_ab ^= 1u; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&r.empty_in[si]);
        if (++si == p.s_in) { si = 0; ph_in ^= 1u; }
        continue;
      }
      if (MODE == LD_DW && p.n_in > 1) load_dw(k0);
      for (int tap = 0; tap < p.taps; ++tap) {
        mbar_wait_spin(&r.empty_ab[sa], ph_ab ^ 1u, p.err_flag);
        if (t == 0) dbg_stamp(p, 0, dbg_kb, 0);
        const uint32_t hi = ab_base + (uint32_t)(sa * r.ab_bytes);
        const uint32_t lo = hi + A_BYTES;
        if (active) {
          float4 acc[PXT];
#pragma unroll
          for (int j = 0; j < PXT; ++j) acc[j] = zero4();
          if (kv) {
            if (MODE == LD_DW) {
#pragma unroll
              for (int j = 0; j < PXT; ++j) acc[j] = b4;
#pragma unroll
              for (int rr = 0; rr < 3; ++rr) {
                const int iy = y * STRIDE - 1 + rr;
                if (iy >= 0 && iy < p.hin) {      // rows outside the image hold a neighbour image's data
                  const uint32_t rb = box + (uint32_t)(rr * row_pitch);
                  float4 v[NV];
#pragma unroll
                  for (int c = 0; c < NV; ++c) v[c] = lds4(rb + c * PP);
#pragma unroll
                  for (int j = 0; j < PXT; ++j)
#pragma unroll
                    for (int s = 0; s < 3; ++s) fma4(acc[j], v[j * STRIDE + s], w[rr * 3 + s]);
                }
              }
#pragma unroll
              for (int j = 0; j < PXT; ++j) {
                acc[j].x = fmaxf(acc[j].x, 0.f); acc[j].y = fmaxf(acc[j].y, 0.f);
                acc[j].z = fmaxf(acc[j].z, 0.f); acc[j].w = fmaxf(acc[j].w, 0.f);
              }
            } else if (MODE == LD_IM2COL) {
              const int dy = tap / 3, dx = tap - dy * 3;
              const int iy = y * STRIDE - 1 + dy;
              if (iy >= 0 && iy < p.hin) {
                const uint32_t rb = box + (uint32_t)(dy * row_pitch + dx * PP);
#pragma unroll
                for (int j = 0; j < PXT; ++j) acc[j] = lds4(rb + j * STRIDE * PP);
              }
            } else {
#pragma unroll
              for (int j = 0; j < PXT; ++j) acc[j] = lds4(box + j * PP);
            }
          }
#pragma unroll
          for (int j = 0; j < PXT; ++j) split_store(hi, lo, soff[j], acc[j]);
        }
        // generic-proxy stores -> visible to the tensor core (async proxy), then publish
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&r.full_ab[sa]);
        if (t == 0) dbg_stamp(p, 0, dbg_kb++, 1);
        if (++sa == p.s_ab) { sa = 0; ph_ab ^= 1u; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.empty_in[si]);   // every lane is done reading the box
      if (++si == p.s_in) { si = 0; ph_in ^= 1u; }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
sep_gemm_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmInTail,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ SepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = p.nt * 128;
  Rings r;
  r.ab_bytes = 2 * A_BYTES + 2 * b_bytes;
  r.in = smem;
  r.ab = smem + p.s_in * p.in_bytes;
  r.full_in = reinterpret_cast<uint64_t*>(r.ab + p.s_ab * r.ab_bytes);
  r.empty_in = r.full_in + MAX_STAGES;
  r.full_ab = r.empty_in + MAX_STAGES;
  r.empty_ab = r.full_ab + MAX_STAGES;
  uint64_t* tfull = r.empty_ab + MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [npad_total]
  float* s_dw = s_bias + p.npad_total;                       // [10][cin]: 9 depthwise taps + bias

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.npad_total; i += THREADS) s_bias[i] = p.bias[i];
  if (p.mode == LD_DW) {
    for (int i = threadIdx.x; i < 9 * p.cin; i += THREADS) s_dw[i] = p.dw_w[i];
    for (int i = threadIdx.x; i < p.cin; i += THREADS) s_dw[9 * p.cin + i] = p.dw_b[i];
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&r.full_in[s], 1);
      mbar_init(&r.empty_in[s], CV_WARPS);
      mbar_init(&r.full_ab[s], CV_WARPS + 1);   // converter warps + the weight TMA's expect_tx arrive
      mbar_init(&r.empty_ab[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::prefetch_tmap(&tmIn);
    tc::prefetch_tmap(&tmInTail);
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
  const int total_tiles = p.num_m_tiles << p.n_shift;
  const int nkb = p.n_in * p.taps;

  if (warp == 0) {
    // ------------------------------------------------------------- input-box TMA
    if (lane == 0) {
      const uint32_t box_bytes = (uint32_t)(p.bh * p.bw * p.kc * 4);
      int s = 0;
      uint32_t ph = 0, dbg_box = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile);
        const int x0 = tc.txi * p.tw;
        for (int i = 0; i < p.n_in; ++i) {
          mbar_wait(&r.empty_in[s], ph ^ 1u, p.err_flag);
          dbg_stamp(p, 3, dbg_box++, 0);
          if (p.tail8 && i == p.n_in - 1) {
            mbar_expect_tx(&r.full_in[s], box_bytes >> 2);
            tma_load_3d(r.in + s * p.in_bytes, &tmInTail, &r.full_in[s], i * p.kc, x0 * p.stride - p.halo,
                        tc.gy0 * p.stride - p.halo);
            if (++s == p.s_in) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(&r.full_in[s], box_bytes);
          if (p.in_merged)
            tma_load_2d(r.in + s * p.in_bytes, &tmIn, &r.full_in[s], (x0 * p.stride - p.halo) * 8,
                        tc.gy0 * p.stride - p.halo);
          else
            tma_load_3d(r.in + s * p.in_bytes, &tmIn, &r.full_in[s], i * p.kc, x0 * p.stride - p.halo,
                        tc.gy0 * p.stride - p.halo);
          if (++s == p.s_in) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------- weight TMA
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, dbg_kb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile & (p.n_tiles_n - 1)) * p.nt;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&r.empty_ab[s], ph ^ 1u, p.err_flag);
          dbg_stamp(p, 2, dbg_kb++, 0);
          mbar_expect_tx(&r.full_ab[s], 2u * (uint32_t)b_bytes);
          uint8_t* sb = r.ab + s * r.ab_bytes + 2 * A_BYTES;
          tma_load_2d(sb, &tmW, &r.full_ab[s], kb * 32, n0);
          tma_load_2d(sb + b_bytes, &tmW, &r.full_ab[s], kb * 32, p.npad_total + n0);
          if (++s == p.s_ab) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    // The whole warp runs this loop in uniform control flow (waits, counters and descriptor
    // arithmetic stay on the uniform datapath); one elected lane issues the MMAs and commits.
    // A shared-memory descriptor is (address >> 4) in its low 14 bits under constant upper
    // bits, so stepping stages / K steps is a 32-bit add on the low word.
    {
      const uint32_t idesc = make_idesc_tf32(TM, p.nt);
      const uint64_t d0 = make_smem_desc(r.ab);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;                         // A_hi tile of stage 0
      const uint32_t stage_step = (uint32_t)r.ab_bytes >> 4;
      const uint32_t off_alo = (uint32_t)A_BYTES >> 4, off_bhi = (uint32_t)(2 * A_BYTES) >> 4;
      const uint32_t off_blo = (uint32_t)(2 * A_BYTES + b_bytes) >> 4;
      const uint32_t nt = (uint32_t)p.nt;
      const bool two_big = p.nbig == 2;
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      uint32_t tt = 0, dbg_kb = 0;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t slot_cols = (uint32_t)((p.nbig + 1) * p.nt);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tt) {
        const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
        const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
        mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d_big = tmem_base + acc * slot_cols;
        const uint32_t d_small = d_big + (uint32_t)p.nbig * nt;
        uint32_t ks_total = 0;
        int tap = 0, crem = p.cin;   // channels left from this K block's box on
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_spin(&r.full_ab[s], ph, p.err_flag);
          if (lane == 0) dbg_stamp(p, 1, dbg_kb, 0);
          tc_fence_after();
          const uint32_t ahi = a_lo0 + (uint32_t)s * stage_step;
          const int ksteps = (min(p.kc, crem) + 7) >> 3;               // K = 8 tf32 per MMA
          if (++tap == p.taps) { tap = 0; crem -= p.kc; }
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < ksteps) {
                const uint32_t o = (uint32_t)(k * 2);   // +32 bytes inside the 128-byte swizzle row
                const uint32_t kt = ks_total + (uint32_t)k;
                mma_tf32(d_small, desc(ahi + off_alo + o), desc(ahi + off_bhi + o), idesc, kt ? 1u : 0u);
                mma_tf32(d_small, desc(ahi + o), desc(ahi + off_blo + o), idesc, 1u);
                const uint32_t which = two_big ? (kt & 1u) : 0u;
                mma_tf32(d_big + which * nt, desc(ahi + o), desc(ahi + off_bhi + o), idesc, kt >= (uint32_t)p.nbig ? 1u : 0u);
              }
            }
            tc_commit(&r.empty_ab[s]);
          }
          __syncwarp();
          ks_total += (uint32_t)ksteps;
          if (lane == 0) dbg_stamp(p, 1, dbg_kb, 1);
          ++dbg_kb;
          if (++s == p.s_ab) { s = 0; ph ^= 1u; }
        }
        if (tc::elect_one()) tc_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------- epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int ty = row / p.tw, tx = row - ty * p.tw;
    uint32_t tt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tt) {
      const TileCoord tc = tile_coord(p, tile);
      const int n0 = tc.n_tile * p.nt;
      const int gy = tc.gy0 + ty;                            // merged output row = n * hout + y
      const int ox = tc.txi * p.tw + tx;
      const bool valid = ty < p.th && gy < p.rows_out;
      const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      if (warp == 2 && lane == 0) dbg_stamp(p, 4, tt, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (uint32_t)((p.nbig + 1) * p.nt) + ((uint32_t)(q * 32) << 16);
      const uint32_t small_off = (uint32_t)(p.nbig * p.nt);
      const size_t pix = (size_t)gy * p.wout + ox;           // NHWC pixel index over the batch
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
            f[i] = __uint_as_float(v0[i]) + s_bias[i];
            f[16 + i] = __uint_as_float(v1[i]) + s_bias[16 + i];
          }
          float2 sc;
          sc.x = 1.0f / (1.0f + expf(-f[0]));
          sc.y = 1.0f / (1.0f + expf(-f[1]));
          *reinterpret_cast<float2*>(p.score + pix * 2) = sc;
          float4* bb = reinterpret_cast<float4*>(p.bbox + pix * 8);
          bb[0] = make_float4(f[2], f[3], f[4], f[5]);
          bb[1] = make_float4(f[6], f[7], f[8], f[9]);
          float4* kp = reinterpret_cast<float4*>(p.kps + pix * 20);
#pragma unroll
          for (int i = 0; i < 5; ++i) kp[i] = make_float4(f[10 + 4 * i], f[11 + 4 * i], f[12 + 4 * i], f[13 + 4 * i]);
        }
      } else {
        const size_t pix_off = valid ? pix * p.cout : 0;
        const float* up = nullptr;
        if (p.add_up && valid) {
          const int n = fastdiv(gy, p.hout_magic), oy = gy - n * p.hout;
          up = p.add_up + ((size_t)(n * (p.hout >> 1) + (oy >> 1)) * (p.wout >> 1) + (ox >> 1)) * p.cout;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
          uint32_t v[16];
          ld_sum16(taddr + c0, small_off, p.nbig, (uint32_t)p.nt, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = n0 + c0 + 4 * j;
              if (c < p.cout) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c);
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
      if (warp == 2 && lane == 0) dbg_stamp(p, 4, tt, 1);
    }
  } else if (warp >= FIRST_CV_WARP) {
    // ------------------------------------------------------------- A-operand converters
    const int t = threadIdx.x - FIRST_CV_WARP * 32;
    if (p.mode == LD_DW) {
      if (p.kc == 32) converter_loop<32, 1, LD_DW>(p, r, s_dw, t, lane);
      else if (p.stride == 1) converter_loop<16, 1, LD_DW>(p, r, s_dw, t, lane);
      else converter_loop<16, 2, LD_DW>(p, r, s_dw, t, lane);
    } else if (p.mode == LD_IM2COL) {
      if (p.kc == 32) converter_loop<32, 1, LD_IM2COL>(p, r, s_dw, t, lane);
      else if (p.stride == 1) converter_loop<16, 1, LD_IM2COL>(p, r, s_dw, t, lane);
      else converter_loop<16, 2, LD_IM2COL>(p, r, s_dw, t, lane);
    } else {
      converter_loop<32, 1, LD_PW>(p, r, s_dw, t, lane);
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
// Dense 3x3 stride-1 convolutions (fpn / pafpn 16 -> 16, the fused 64 -> 30 head conv) in "halo
// mode": the tile is a th x tw patch whose tile rows use the pitch tw2 = tw + 2 of its own halo
// box, so the (th+2) x tw2 input box is split into tf32 hi/lo ONCE, as one operand block of
// box-pixel rows, and the nine taps are row-shifted shared-memory descriptors
// (start + (dy*tw2 + dx) rows) on it -- instead of nine im2col copies with nine hand-offs.
// Tile row ty*tw2 + tx is output pixel (ty, tx) if ty < th and tx < tw, a dead row otherwise
// (7 x 16 of 128 rows live at 80^2, 5 x 20 at 40^2 / 20^2).  Weights stream through their own
// ring (or stay resident when all 9 * n_in blocks fit); roles and accumulators as in
// sep_gemm_kernel.
struct C3Params {
  float* out;
  const float* bias;
  float* score;
  float* bbox;
  float* kps;
  int cin, cout, h, w;        // square maps: h == w
  int n_img, epi;
  int kc, n_in, nt, npad_total;
  int th, tw, tw2, tiles_x, tiles_y;
  uint32_t tiles_per_img_magic, tiles_x_magic, tw2_magic, chunk_magic;
  int num_tiles;
  int box_px;                 // (th + 2) * tw2
  int in_bytes, a_rows;       // bytes per input stage (1024-rounded), rows per A operand block
  int s_in, s_a, s_b, b_resident;
  int tmem_cols, acc_stages;
  int* err_flag;
};

constexpr int C3_MAX_B = 20;

__global__ void __launch_bounds__(THREADS, 1)
conv3_halo_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ C3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_blk = p.a_rows * 128;            // one hi (or lo) operand block
  const int b_bytes = p.nt * 128;
  uint8_t* s_inr = smem;
  uint8_t* s_a = s_inr + p.s_in * p.in_bytes;
  uint8_t* s_b = s_a + p.s_a * 2 * a_blk;
  uint64_t* full_in = reinterpret_cast<uint64_t*>(s_b + p.s_b * 2 * b_bytes);
  uint64_t* empty_in = full_in + MAX_STAGES;
  uint64_t* full_a = empty_in + MAX_STAGES;
  uint64_t* empty_a = full_a + MAX_STAGES;
  uint64_t* full_b = empty_a + MAX_STAGES;
  uint64_t* empty_b = full_b + C3_MAX_B;
  uint64_t* tfull = empty_b + C3_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.npad_total; i += THREADS) s_bias[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full_in[s], 1);
      mbar_init(&empty_in[s], CV_WARPS);
      mbar_init(&full_a[s], CV_WARPS);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < C3_MAX_B; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::prefetch_tmap(&tmIn);
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
  const int nkb = p.n_in * 9;

  if (warp == 0) {
    // ------------------------------------------------------------- input-box TMA
    if (lane == 0) {
      const uint32_t box_bytes = (uint32_t)(p.box_px * p.kc * 4);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int img = fastdiv(tile, p.tiles_per_img_magic);
        const int rem = tile - img * p.tiles_x * p.tiles_y;
        const int tyi = fastdiv(rem, p.tiles_x_magic);
        const int y0 = tyi * p.th, x0 = (rem - tyi * p.tiles_x) * p.tw;
        for (int i = 0; i < p.n_in; ++i) {
          mbar_wait(&empty_in[s], ph ^ 1u, p.err_flag);
          mbar_expect_tx(&full_in[s], box_bytes);
          tma_load_3d(s_inr + s * p.in_bytes, &tmIn, &full_in[s], i * p.kc, x0 - 1, img * p.h + y0 - 1);
          if (++s == p.s_in) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------- weight TMA
    if (lane == 0) {
      if (p.b_resident) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_expect_tx(&full_b[kb], 2u * (uint32_t)b_bytes);
          uint8_t* sb = s_b + kb * 2 * b_bytes;
          tma_load_2d(sb, &tmW, &full_b[kb], kb * 32, 0);
          tma_load_2d(sb + b_bytes, &tmW, &full_b[kb], kb * 32, p.npad_total);
        }
      } else {
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_b[s], ph ^ 1u, p.err_flag);
            mbar_expect_tx(&full_b[s], 2u * (uint32_t)b_bytes);
            uint8_t* sb = s_b + s * 2 * b_bytes;
            tma_load_2d(sb, &tmW, &full_b[s], kb * 32, 0);
            tma_load_2d(sb + b_bytes, &tmW, &full_b[s], kb * 32, p.npad_total);
            if (++s == p.s_b) { s = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (warp-uniform)
    const uint32_t idesc = make_idesc_tf32(TM, p.nt), idesc2 = make_idesc_tf32(TM, 2 * p.nt);
    const uint64_t d0 = make_smem_desc(s_a);
    const uint32_t desc_hi = (uint32_t)(d0 >> 32);
    const uint32_t a_lo0 = (uint32_t)d0;
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(s_b);
    const uint32_t a_step = (uint32_t)(2 * a_blk) >> 4, a_lo_off = (uint32_t)a_blk >> 4;
    const uint32_t b_step = (uint32_t)(2 * b_bytes) >> 4;   // a stage = w_hi tile + w_lo tile, contiguous
    const uint32_t nt = (uint32_t)p.nt;
    const int ksteps = p.kc >> 3;
    uint32_t tt = 0;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tt) {
      const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
      mbar_wait(&tempty[acc], acc_phase ^ 1u, p.err_flag);
      tc_fence_after();
      // three independent accumulator regions ([hh0 | hl0], [hh1 | hl1], lh): consecutive MMAs never
      // target the same one, so the accumulate-dependency latency of these short MMAs is hidden
      const uint32_t d_big = tmem_base + acc * 5u * nt;
      int kb = 0;
      for (int i = 0; i < p.n_in; ++i) {
        mbar_wait_spin(&full_a[sa], pa, p.err_flag);
        tc_fence_after();
        const uint32_t ahi0 = a_lo0 + (uint32_t)sa * a_step;
        // one elected lane runs all nine taps of the box (waits included): tap offsets are
        // compile-time constants and there is no per-tap warp re-convergence
        if (tc::elect_one()) {
          int sb_l = sb;
          uint32_t pb_l = pb;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int bslot = p.b_resident ? kb + tap : sb_l;
            if (!p.b_resident || tt == 0) {
              mbar_wait_spin(&full_b[bslot], p.b_resident ? 0u : pb_l, p.err_flag);
              tc_fence_after();
            }
            const uint32_t ahi = ahi0 + (uint32_t)((tap / 3) * p.tw2 + (tap % 3)) * 8u;   // 128-byte rows = 8 units
            const uint32_t bhi = b_lo0 + (uint32_t)bslot * b_step;
            const uint32_t alo = ahi + a_lo_off;                                  // w_lo follows w_hi in the stage
            if (ksteps == 4) {
              if (tap == 0 && i == 0) c3_issue_tap<4, true>(d_big, nt, desc_hi, ahi, alo, bhi, idesc, idesc2);
              else c3_issue_tap<4, false>(d_big, nt, desc_hi, ahi, alo, bhi, idesc, idesc2);
            } else {
              if (tap == 0 && i == 0) c3_issue_tap<2, true>(d_big, nt, desc_hi, ahi, alo, bhi, idesc, idesc2);
              else c3_issue_tap<2, false>(d_big, nt, desc_hi, ahi, alo, bhi, idesc, idesc2);
            }
            if (!p.b_resident) {
              tc_commit(&empty_b[sb_l]);
              if (++sb_l == p.s_b) { sb_l = 0; pb_l ^= 1u; }
            }
          }
        }
        __syncwarp();
        kb += 9;
        if (!p.b_resident) {   // keep the ring position uniform across the warp
          sb += 9;
          while (sb >= p.s_b) { sb -= p.s_b; pb ^= 1u; }
        }
        if (tc::elect_one()) tc_commit(&empty_a[sa]);
        __syncwarp();
        if (++sa == p.s_a) { sa = 0; pa ^= 1u; }
      }
      if (tc::elect_one()) tc_commit(&tfull[acc]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = fastdiv(row, p.tw2_magic), tx = row - ty * p.tw2;
    uint32_t tt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tt) {
      const int img = fastdiv(tile, p.tiles_per_img_magic);
      const int rem = tile - img * p.tiles_x * p.tiles_y;
      const int tyi = fastdiv(rem, p.tiles_x_magic);
      const int y = tyi * p.th + ty, x = (rem - tyi * p.tiles_x) * p.tw + tx;
      const bool valid = ty < p.th && tx < p.tw && y < p.h;
      const uint32_t acc = p.acc_stages == 2 ? (tt & 1u) : 0u;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((tt >> 1) & 1u) : (tt & 1u);
      mbar_wait(&tfull[acc], acc_phase, p.err_flag);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 5u * (uint32_t)p.nt + ((uint32_t)(q * 32) << 16);
      const size_t pix = ((size_t)img * p.h + y) * p.w + x;
      if (p.epi == EPI_HEAD) {
        uint32_t v0[16], v1[16];
        ld_sum16x5(taddr, (uint32_t)p.nt, v0);
        ld_sum16x5(taddr + 16, (uint32_t)p.nt, v1);
        if (valid) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = __uint_as_float(v0[i]) + s_bias[i];
            f[16 + i] = __uint_as_float(v1[i]) + s_bias[16 + i];
          }
          float2 sc;
          sc.x = 1.0f / (1.0f + expf(-f[0]));
          sc.y = 1.0f / (1.0f + expf(-f[1]));
          *reinterpret_cast<float2*>(p.score + pix * 2) = sc;
          float4* bb = reinterpret_cast<float4*>(p.bbox + pix * 8);
          bb[0] = make_float4(f[2], f[3], f[4], f[5]);
          bb[1] = make_float4(f[6], f[7], f[8], f[9]);
          float4* kp = reinterpret_cast<float4*>(p.kps + pix * 20);
#pragma unroll
          for (int i = 0; i < 5; ++i) kp[i] = make_float4(f[10 + 4 * i], f[11 + 4 * i], f[12 + 4 * i], f[13 + 4 * i]);
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
          uint32_t v[16];
          ld_sum16x5(taddr + c0, (uint32_t)p.nt, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 4 * j;
              if (c < p.cout) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c);
                *reinterpret_cast<float4*>(p.out + pix * p.cout + c) =
                    make_float4(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y,
                                __uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  } else if (warp >= FIRST_CV_WARP) {
    // ------------------------------------------------------------- box -> hi/lo operand block
    const int t = threadIdx.x - FIRST_CV_WARP * 32;
    const int ch = p.kc >> 2;                 // 16-byte chunks per box pixel
    const int pp = p.kc * 4;                  // box pixel pitch (bytes)
    const int items = p.box_px * ch;
    const uint32_t in_base = smem_u32(s_inr), a_base = smem_u32(s_a);
    int si = 0, sa = 0;
    uint32_t ph_in = 0, ph_a = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = fastdiv(tile, p.tiles_per_img_magic);
      const int rem = tile - img * p.tiles_x * p.tiles_y;
      const int y0 = fastdiv(rem, p.tiles_x_magic) * p.th;
      (void)img;
      for (int i = 0; i < p.n_in; ++i) {
        mbar_wait_spin(&full_in[si], ph_in, p.err_flag);
        mbar_wait_spin(&empty_a[sa], ph_a ^ 1u, p.err_flag);
        const uint32_t box = in_base + (uint32_t)(si * p.in_bytes);
        const uint32_t hi = a_base + (uint32_t)(sa * 2 * a_blk);
        const uint32_t lo = hi + (uint32_t)a_blk;
        for (int it = t; it < items; it += CV_WARPS * 32) {
          const int px = fastdiv(it, p.chunk_magic);     // box pixel = operand row
          const int chunk = it - px * ch;
          const int r = fastdiv(px, p.tw2_magic);         // box row; image row = y0 - 1 + r
          const int iy = y0 - 1 + r;
          float4 v = zero4();
          // rows outside the image hold a neighbour image's data (merged-row tensor map): zero them
          if (iy >= 0 && iy < p.h) v = lds4(box + (uint32_t)(px * pp + chunk * 16));
          split_store(hi, lo, sw128(px, chunk), v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&full_a[sa]);
          mbar_arrive(&empty_in[si]);
        }
        if (++sa == p.s_a) { sa = 0; ph_a ^= 1u; }
        if (++si == p.s_in) { si = 0; ph_in ^= 1u; }
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
// One thread = 4 horizontally adjacent output pixels x 16 channels: per (channel, row) one
// 16-byte load brings input columns 2*ox .. 2*ox+7 and one 2-byte load column 2*ox-1; the
// weights are broadcast from shared memory as float4 (1 LDS.128 per 16 FMAs); the thread
// stores 256 contiguous bytes of NHWC output.
__global__ void __launch_bounds__(128)
stem_conv_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                 const float* __restrict__ b, int n_img) {
  __shared__ __align__(16) float sw[27 * 16];
  __shared__ __align__(16) float sb[16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 16) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  constexpr int HO = DET / 2, QW = HO / 4;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)n_img * HO * QW) return;
  const int n = (int)(gid / (HO * QW));
  const int rem = (int)(gid - (size_t)n * HO * QW);
  const int oy = rem / QW, ox0 = (rem - oy * QW) * 4;
  float acc[4][16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[j][i] = sb[i];
  const __nv_bfloat16* ip = in + (size_t)n * 3 * DET * DET;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 - 1 + r;
      float v[9];   // input columns 2*ox0-1 .. 2*ox0+7
#pragma unroll
      for (int i = 0; i < 9; ++i) v[i] = 0.f;
      if (iy >= 0 && iy < DET) {
        const __nv_bfloat16* rp = ip + ((size_t)c * DET + iy) * DET + 2 * ox0;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp));
        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[1 + 2 * i] = __uint_as_float(qq[i] << 16);
          v[2 + 2 * i] = __uint_as_float(qq[i] & 0xffff0000u);
        }
        if (ox0 > 0) v[0] = __bfloat162float(rp[-1]);
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4* wp = reinterpret_cast<const float4*>(&sw[((c * 3 + r) * 3 + s) * 16]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w4 = wp[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = v[2 * j + s];
            acc[j][4 * i] = fmaf(x, w4.x, acc[j][4 * i]);
            acc[j][4 * i + 1] = fmaf(x, w4.y, acc[j][4 * i + 1]);
            acc[j][4 * i + 2] = fmaf(x, w4.z, acc[j][4 * i + 2]);
            acc[j][4 * i + 3] = fmaf(x, w4.w, acc[j][4 * i + 3]);
          }
        }
      }
    }
  float4* op = reinterpret_cast<float4*>(out + (((size_t)n * HO + oy) * HO + ox0) * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      op[j * 4 + i] = make_float4(fmaxf(acc[j][4 * i], 0.f), fmaxf(acc[j][4 * i + 1], 0.f),
                                  fmaxf(acc[j][4 * i + 2], 0.f), fmaxf(acc[j][4 * i + 3], 0.f));
}


// ---------------------------------------------------------------------------------------
// The same stem on warp-level MMA (the default).  The CUDA-core kernel above is instruction bound
// (432 FMAs per pixel) at 2.1 TB/s; the layer's bound is the 6.5 MB / frame write.  Here a warp
// computes 16 pixels x 16 channels per step as an m16n8k16 bf16 GEMM with K = 27 (+ a bias row,
// padded to 32).  K1's inputs (v - 127.5) / 128 are exact in bf16 and the fp32 weights enter as
// three bf16 pieces (24 bits: exact), so every product is exact in fp32 and the result differs from
// the FMA chain only by fp32 summation order.  A block walks SM_GROUPS groups of 4 output rows x half
// a frame; each group's 9 input rows x 3 planes are staged in shared memory by cp.async, double
// buffered so the next group lands while this one computes; the C fragments pass through a padded
// per-warp tile so that every store instruction writes 512 contiguous bytes (8 pixels x 64 B).
constexpr int SM_ROWS = 9, SM_COLS = DET / 2, SM_PITCH = SM_COLS + 16;   // strip: half a frame wide; ix0 - 1 at index 7
constexpr int SM_STAGE_PITCH = 96;                                // bytes per staged pixel (64 + 32: conflict-free)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

constexpr int SM_GROUPS = 4;                                      // 4-row groups per block (double-buffered strips)
constexpr int SM_STRIP_ELEMS = 3 * SM_ROWS * SM_PITCH;
constexpr int SM_SMEM_BYTES = 2 * SM_STRIP_ELEMS * 2 + 4 * 16 * SM_STAGE_PITCH;

__global__ void __launch_bounds__(128)
stem_conv_mma_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, const uint2* __restrict__ bfrag) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  uint16_t* strips = reinterpret_cast<uint16_t*>(sm_raw);                        // [2][3][SM_ROWS][SM_PITCH]
  uint8_t* stage_base = sm_raw + 2 * SM_STRIP_ELEMS * 2;                         // [4 warps][16 px][96 B]
  constexpr int HO = DET / 2, BLOCKS_PER_IMG = 2 * HO / (4 * SM_GROUPS);
  const int n = blockIdx.x / BLOCKS_PER_IMG;
  const int bi = blockIdx.x - n * BLOCKS_PER_IMG;
  const int oyb = (bi >> 1) * (4 * SM_GROUPS);
  const int ix0 = (bi & 1) * SM_COLS, ox0 = ix0 / 2;   // this block's half of the frame
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // strip[c][row][8 + ix - ix0] = input (c, 2*oy0 - 1 + row, ix); the 8 elements before ix0 are zero for the left half.
  // Filled with 16-byte cp.async so that the next group's rows land while this group computes.
  const __nv_bfloat16* ip = in + (size_t)n * 3 * DET * DET;
  auto fill = [&](int buf, int oy0) {
    uint16_t* sp = strips + buf * SM_STRIP_ELEMS;
    for (int i = threadIdx.x; i < 3 * SM_ROWS * (SM_COLS / 8 + 1); i += 128) {
      const int cr = i / (SM_COLS / 8 + 1), q = i - cr * (SM_COLS / 8 + 1) - 1;   // q = -1: the 8 columns before ix0
      const int c = cr / SM_ROWS, row = cr - c * SM_ROWS, iy = 2 * oy0 - 1 + row;
      uint16_t* dst = sp + (size_t)cr * SM_PITCH + 8 + 8 * q;
      if (iy >= 0 && ix0 + 8 * q >= 0) {
        const uint32_t d = smem_u32(dst);
        const void* src = ip + ((size_t)c * DET + iy) * DET + ix0 + 8 * q;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill(0, oyb);
  // per-thread constants: strip offsets of its 8 k indices (k = (c*3 + r)*3 + s), weight fragments
  int koff[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int k = (q >> 2) * 16 + ((q >> 1) & 1) * 8 + 2 * t + (q & 1);
    const int c = k / 9, r = (k - c * 9) / 3, sx = k - c * 9 - r * 3;
    koff[q] = k < 27 ? (c * SM_ROWS + r) * SM_PITCH + 7 + sx : 0;
  }
  uint2 bf[2][2][3];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 3; ++h) bf[ks][j][h] = __ldg(bfrag + ((ks * 2 + j) * 3 + h) * 32 + lane);
  uint8_t* st = stage_base + warp * 16 * SM_STAGE_PITCH;
  // read-back mapping: a quarter warp takes pixels r and r + 2 (distinct banks at a 96-byte pitch)
  const int rq = ((lane >> 4) << 2) + ((lane >> 3) & 1) + (((lane >> 2) & 1) << 1);
#pragma unroll 1
  for (int gr = 0; gr < SM_GROUPS; ++gr) {
    const int oy0 = oyb + 4 * gr;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // this group's strip is complete; the other buffer is free
    if (gr + 1 < SM_GROUPS) fill((gr + 1) & 1, oy0 + 4);
    const uint16_t* strip = strips + (gr & 1) * SM_STRIP_ELEMS;
    const int oy = oy0 + warp;
    float* orow = out + (((size_t)n * HO + oy) * HO + ox0) * 16;
#pragma unroll 1
    for (int m = 0; m < HO / 32; ++m) {
      const uint16_t* s0 = strip + (2 * warp) * SM_PITCH + 2 * (16 * m + g);
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                     // a0:(g,k lo) a1:(g+8,k lo) a2:(g,k hi) a3:(g+8,k hi)
          const uint16_t* sp = s0 + (q & 1) * 16;
          const int ko = ks * 4 + (q >> 1) * 2;
          a[q] = (uint32_t)sp[koff[ko]] | ((uint32_t)sp[koff[ko + 1]] << 16);
        }
        if (ks == 1 && t == 1) {                          // k = 27 carries the bias: A = 1.0
          a[2] = (a[2] & 0xffffu) | 0x3f800000u;
          a[3] = (a[3] & 0xffffu) | 0x3f800000u;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 2; h >= 0; --h) mma_bf16_16816(acc[j], a, bf[ks][j][h]);   // small pieces first
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        *reinterpret_cast<float2*>(st + g * SM_STAGE_PITCH + 32 * j + 8 * t) =
            make_float2(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f));
        *reinterpret_cast<float2*>(st + (g + 8) * SM_STAGE_PITCH + 32 * j + 8 * t) =
            make_float2(fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
      }
      __syncwarp();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = pass * 8 + rq;
        const float4 v = *reinterpret_cast<const float4*>(st + r * SM_STAGE_PITCH + 16 * t);
        *reinterpret_cast<float4*>(orow + (size_t)(16 * m + r) * 16 + 4 * t) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// b0 (depthwise 3x3 + ReLU -> 1x1 16 -> 16 + ReLU on the 320 x 320 x 16 stem output) on warp-level MMA.
// K = N = 16: the tcgen05 pipeline above spends its time on hand-offs (348 us for a 0.84 GB
// stream).  Here a warp owns 16 pixels per step: each thread computes the depthwise outputs of
// 2 pixels x 4 channels on CUDA cores straight into the A fragment of an m16n8k8 tf32 MMA (the
// GEMM's k slots are a permutation of the channels so that those 4 channels are one float4 of the
// NHWC input), splits them hi/lo like the converter warps do, and the three-term product runs with
// separate small / big accumulators.  The C fragment of a quad is one 32-byte sector per pixel, so
// it is stored directly.  Block = 8 warps = 8 output rows x 64 pixels per group; the 10 x 66 pixel
// input tile arrives by cp.async, double buffered over B0_GROUPS groups.
constexpr int B0_TW = 64, B0_ROWS = 8, B0_TP = B0_TW + 2, B0_TR = B0_ROWS + 2, B0_GROUPS = 4;
constexpr int B0_TILE_FLOATS = B0_TR * B0_TP * 16;
constexpr int B0_SMEM_BYTES = 2 * B0_TILE_FLOATS * 4;

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], float2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
}

__global__ void __launch_bounds__(256, 2)
b0_dwpw_mma_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ dw_w,
                   const float* __restrict__ dw_b, const float2* __restrict__ bfrag, const float* __restrict__ pw_b,
                   int hw) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* tiles = reinterpret_cast<float*>(sm_raw);                    // [2][B0_TR][B0_TP][16]
  const int strips = hw / B0_TW, row_blocks = hw / (B0_ROWS * B0_GROUPS);
  const int n = blockIdx.x / (strips * row_blocks);
  const int bi = blockIdx.x - n * (strips * row_blocks);
  const int x0 = (bi % strips) * B0_TW, oyb = (bi / strips) * (B0_ROWS * B0_GROUPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* ip = in + (size_t)n * hw * hw * 16;
  // tile[row][px][16] = input (oy0 - 1 + row, x0 - 1 + px); outside the frame: zero (the dw conv's padding)
  auto fill = [&](int buf, int oy0) {
    float* tp = tiles + buf * B0_TILE_FLOATS;
    for (int i = threadIdx.x; i < B0_TR * B0_TP * 4; i += 256) {
      const int row = i / (B0_TP * 4), rem = i - row * (B0_TP * 4);
      const int iy = oy0 - 1 + row, ix = x0 - 1 + (rem >> 2);
      float* dst = tp + (size_t)i * 4;
      if (iy >= 0 && iy < hw && ix >= 0 && ix < hw) {
        const uint32_t d = smem_u32(dst);
        const void* src = ip + ((size_t)iy * hw + ix) * 16 + (rem & 3) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill(0, oyb);
  // per-thread constants: depthwise taps / bias of channels 4t .. 4t+3, 1x1 weight fragments, 1x1 bias
  float4 wd[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) wd[q] = ldg4(dw_w + q * 16 + 4 * t);
  const float4 bd = ldg4(dw_b + 4 * t);
  float2 bw[2][2][2];                                   // [k step][n tile][hi / lo]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) bw[ks][j][h] = __ldg(bfrag + ((ks * 2 + j) * 2 + h) * 32 + lane);
  float2 pb[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) pb[j] = __ldg(reinterpret_cast<const float2*>(pw_b + 8 * j + 2 * t));
#pragma unroll 1
  for (int gr = 0; gr < B0_GROUPS; ++gr) {
    const int oy0 = oyb + B0_ROWS * gr;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                    // this group's tile is complete; the other buffer is free
    if (gr + 1 < B0_GROUPS) fill((gr + 1) & 1, oy0 + B0_ROWS);
    const uint32_t tile = smem_u32(tiles + (gr & 1) * B0_TILE_FLOATS);
    const int oy = oy0 + warp;
    float* orow = out + (((size_t)n * hw + oy) * hw + x0) * 16;
#pragma unroll 1
    for (int m = 0; m < B0_TW / 16; ++m) {
      // depthwise 3x3 + bias + ReLU of pixels (16m + g) and (16m + g + 8), channels 4t .. 4t+3
      float4 x[2];
#pragma unroll
      for (int hpx = 0; hpx < 2; ++hpx) {
        float4 a = bd;
        const uint32_t base = tile + (uint32_t)(((warp * B0_TP) + 16 * m + g + 8 * hpx) * 64 + 16 * t);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) fma4(a, lds4(base + (uint32_t)((r * B0_TP + dx) * 64)), wd[r * 3 + dx]);
        x[hpx] = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
      }
      const float xs[2][4] = {{x[0].x, x[0].y, x[0].z, x[0].w}, {x[1].x, x[1].y, x[1].z, x[1].w}};
      float accS[2][4], accB[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) { accS[j][i] = 0.f; accB[j][i] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // k slot t <-> channel 4t + 2ks, k slot t + 4 <-> channel 4t + 2ks + 1 (the weights are permuted to match)
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                     // a0:(g, slot t) a1:(g+8, slot t) a2:(g, t+4) a3:(g+8, t+4)
          const float v = xs[q & 1][2 * ks + (q >> 1)];
          const float h = tf32_rna(v);
          ahi[q] = __float_as_uint(h);
          alo[q] = __float_as_uint(tf32_rna(v - h));
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          mma_tf32_1688(accS[j], alo, bw[ks][j][0]);
          mma_tf32_1688(accS[j], ahi, bw[ks][j][1]);
          mma_tf32_1688(accB[j], ahi, bw[ks][j][0]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 lo = make_float2(fmaxf(accB[j][0] + accS[j][0] + pb[j].x, 0.f), fmaxf(accB[j][1] + accS[j][1] + pb[j].y, 0.f));
        const float2 hi = make_float2(fmaxf(accB[j][2] + accS[j][2] + pb[j].x, 0.f), fmaxf(accB[j][3] + accS[j][3] + pb[j].y, 0.f));
        *reinterpret_cast<float2*>(orow + (size_t)(16 * m + g) * 16 + 8 * j + 2 * t) = lo;
        *reinterpret_cast<float2*>(orow + (size_t)(16 * m + g + 8) * 16 + 8 * j + 2 * t) = hi;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Front of the backbone in ONE kernel: stem (3x3 s2, 3 -> 16) -> b0 (dw 3x3 + 1x1 16 -> 16) -> s0.0
// (dw 3x3 s2 + 1x1 16 -> 40).  As three launches these layers stream the 320 x 320 x 16 fp32 maps
// through HBM four times (2.06 GB per 64 frames, 522 us); fused, a block walks down a strip of the
// frame and keeps the stem / b0 rows it still needs in two shared-memory rings, so the only HBM
// traffic is the bf16 input (157 MB) and the 160 x 160 x 40 output (262 MB).
//   unit   = (frame, strip of OW = 40 s0.0 output columns, segment of G groups); group = 4 b0 rows;
//   per group (base b0 row y): [stem rows y+1..y+4] sync [b0 rows y..y+3] sync [s0.0 rows y/2, y/2+1 ->
//              global]; a unit's prologue is a group that computes stem rows y0-3..y0 and b0 rows
//              y0-2, y0-1 only; positions outside the frame are stored as zeros (the next conv's padding);
//   tiles  : an MMA tile is 2 rows x 8 columns (fragment row g = upper pixel, g + 8 = the pixel below),
//              so a thread owns ONE column: the b0 stencil of a 4-row block reads 6 rows x 3 columns
//              (18 float4 loads for 4 pixels instead of 36), and ring rows are warp-uniform, so a stencil
//              address is uniform row base + per-thread column offset;
//   arithmetic: exactly the stem / b0 kernels' (same fragments, same accumulation order), and for
//              s0.0 the b0 scheme with 5 n tiles (three-term tf32 split, separate small / big sums).
// Pixels are 64-byte records (16 channels).  Stem ring: bit 1 of the 16-byte chunk index is XORed with
// bit 1 of the column, which keeps a quarter warp's float4 loads (two adjacent records) conflict-free and
// spreads a half warp's C-fragment stores (4 columns x 4 x 8 bytes) over all banks.  b0 ring: even and odd columns in
// two planes (offset = 32 mod 128 bytes), so the stride-2 stencil of s0.0 also reads adjacent records.
namespace ff {
constexpr int HS = DET / 2, HO = DET / 4;
constexpr int OW = 40;                 // s0.0 output columns per strip
constexpr int NB = 2 * OW + 1;         // b0 columns a strip needs   (2*X0 - 1 .. 2*X0 + 79)
constexpr int NS = 2 * OW + 3;         // stem columns               (2*X0 - 2 .. 2*X0 + 80)
constexpr int CBS = (NS + 7) / 8;      // 8-column blocks of a stem row (11)
constexpr int CBB = (NB + 7) / 8;      // of a b0 row (11)
constexpr int IW = 176;                // input columns staged per strip row: 4*X0 - 8 .. 4*X0 + 167
constexpr int IP = 208;                // strip row pitch (bf16): 104 words = 8 mod 32, so the rows of a filter window
constexpr int IROWS = 9;               // (input rows 2y+1 .. 2y+9 of a group) start 8 banks apart ...
constexpr int IPLANE = IROWS * IP + 32;   // ... and so do the planes (9 * 8 + 16 = 24 mod 32 = row index 3 * 8)
constexpr int STRIP_ELEMS = 3 * IPLANE;
constexpr int SROWS = 6, BROWS = 5;
constexpr int S_ROWB = (NS + 1) * 64;
constexpr int B_PLANE = (OW + 1) * 64 + 32;          // even columns first, then the odd ones
constexpr int B_ROWB = B_PLANE + OW * 64;
constexpr int WARPS = 6, FTHREADS = WARPS * 32;
constexpr int OFF_SRING = 2 * STRIP_ELEMS * 2;
constexpr int OFF_BRING = OFF_SRING + SROWS * S_ROWB;
constexpr int OFF_STEM_BF = OFF_BRING + BROWS * B_ROWB;          // uint2 [2][2][3][32]
constexpr int OFF_B0_BF = OFF_STEM_BF + 2 * 2 * 3 * 32 * 8;      // float2 [2][2][2][32]
constexpr int OFF_S0_BF = OFF_B0_BF + 2 * 2 * 2 * 32 * 8;        // float2 [2][5][2][32]
constexpr int OFF_PB0 = OFF_S0_BF + 2 * 5 * 2 * 32 * 8;          // float [16]
constexpr int OFF_PB1 = OFF_PB0 + 16 * 4;                        // float [40]
constexpr int SMEM_BYTES = OFF_PB1 + 40 * 4;
static_assert(SMEM_BYTES * 2 + 2048 <= 227 * 1024, "two blocks per SM");
static_assert(WARPS >= OW / 8, "one s0.0 tile per warp");
static_assert(OFF_SRING % 16 == 0 && OFF_BRING % 16 == 0 && B_ROWB % 16 == 0 && OFF_STEM_BF % 16 == 0, "alignment");

struct Params {
  const __nv_bfloat16* in;   // [n][3][640][640]
  float* out;                // s0.0 output, fp32 NHWC [n][160][160][40]
  const uint2* stem_bf;
  const float2* b0_bf;
  const float2* s0_bf;
  const float *pb0, *pb1;    // 1x1 biases: b0 [16], s0.0 [40]
  float dw0[10 * 16];        // b0 depthwise taps [9][16] + bias [16]: read through the constant bank (4 distinct
  float dw1[10 * 16];        // addresses per warp), which keeps 20 float4 loads per phase off the shared-memory pipe
  float* tap_stem;           // TAPS: [n][320][320][16] copies of the intermediates (test hook)
  float* tap_b0;
  int groups;                // G: groups per unit (divides 80)
};

__device__ __forceinline__ void sts2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
// byte offset of 16-byte chunk `chunk` of stem-ring column js inside a ring row
__device__ __forceinline__ uint32_t s_off(int js, int chunk) {
  return (uint32_t)(js * 64 + ((chunk ^ (js & 2)) << 4));
}
__device__ __forceinline__ float4 relu4(const float4& a) {
  return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}
// the 1x1 of 16 pixels (fragment rows g: xa, g + 8: xb; this thread's channels 4t .. 4t+3) against NT n tiles:
// three-term tf32 split, cross terms in accS, hi*hi in accB (k slot t <-> channel 4t + 2ks, t + 4 <-> 4t + 2ks + 1)
template <int NT, typename LoadB>
__device__ __forceinline__ void pw_mma(const float4& xa, const float4& xb, float (&accS)[NT][4], float (&accB)[NT][4],
                                       LoadB load_b) {
  const float xs[2][4] = {{xa.x, xa.y, xa.z, xa.w}, {xb.x, xb.y, xb.z, xb.w}};
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { accS[j][q] = 0.f; accB[j][q] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = xs[q & 1][2 * ks + (q >> 1)];
      const float hh = tf32_rna(v);
      ahi[q] = __float_as_uint(hh);
      alo[q] = __float_as_uint(tf32_rna(v - hh));
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float2 bhi = load_b(ks, j, 0), blo = load_b(ks, j, 1);
      mma_tf32_1688(accS[j], alo, bhi);
      mma_tf32_1688(accS[j], ahi, blo);
      mma_tf32_1688(accB[j], ahi, bhi);
    }
  }
}
}  // namespace ff

template <bool TAPS>
__global__ void __launch_bounds__(ff::FTHREADS, 2)
front_fused_kernel(const ff::Params P) {
  using namespace ff;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  uint16_t* strips = reinterpret_cast<uint16_t*>(sm_raw);
  const uint32_t sm0 = smem_u32(sm_raw);
  const uint32_t sring = sm0 + OFF_SRING, bring = sm0 + OFF_BRING;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int segs = (HS / 4) / P.groups;
  int b = blockIdx.x;
  const int seg = b % segs;
  b /= segs;
  const int strip_i = b & 3, n = b >> 2;
  const int X0 = strip_i * OW;
  const int y0 = seg * 4 * P.groups;
  const __nv_bfloat16* ip = P.in + (size_t)n * 3 * DET * DET;

  // input strip of the group with base row y: strip[c][ir][li] = input (c, 2y + 1 + ir, 4*X0 - 8 + li)
  auto fill = [&](int buf, int y) {
    uint16_t* sp = strips + buf * STRIP_ELEMS;
    for (int i = tid; i < 3 * IROWS * (IW / 8); i += FTHREADS) {
      const int cr = i / (IW / 8), q = i - cr * (IW / 8);
      const int c = cr / IROWS, ir = cr - c * IROWS;
      const int iy = 2 * y + 1 + ir, ic = 4 * X0 - 8 + 8 * q;
      uint16_t* dst = sp + c * IPLANE + ir * IP + 8 * q;
      if (iy >= 0 && iy < DET && ic >= 0 && ic < DET) {
        const void* src = ip + ((size_t)c * DET + iy) * DET + ic;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill(0, y0 - 4);
  // constants -> shared memory
  {
    uint32_t* d = reinterpret_cast<uint32_t*>(sm_raw + OFF_STEM_BF);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(P.stem_bf);
    for (int i = tid; i < 2 * 2 * 3 * 32 * 2; i += FTHREADS) d[i] = __ldg(s + i);
    d = reinterpret_cast<uint32_t*>(sm_raw + OFF_B0_BF);
    s = reinterpret_cast<const uint32_t*>(P.b0_bf);
    for (int i = tid; i < 2 * 2 * 2 * 32 * 2; i += FTHREADS) d[i] = __ldg(s + i);
    d = reinterpret_cast<uint32_t*>(sm_raw + OFF_S0_BF);
    s = reinterpret_cast<const uint32_t*>(P.s0_bf);
    for (int i = tid; i < 2 * 5 * 2 * 32 * 2; i += FTHREADS) d[i] = __ldg(s + i);
    float* f = reinterpret_cast<float*>(sm_raw + OFF_PB0);
    if (tid < 16) f[tid] = __ldg(P.pb0 + tid);
    f = reinterpret_cast<float*>(sm_raw + OFF_PB1);
    if (tid < 40) f[tid] = __ldg(P.pb1 + tid);
  }
  // stem: strip offsets of this thread's 8 k indices (k = (c*3 + r)*3 + s; k = 27 is the bias row)
  int koff[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int k = (q >> 2) * 16 + ((q >> 1) & 1) * 8 + 2 * t + (q & 1);
    const int c = k / 9, r = (k - c * 9) / 3, sx = k - c * 9 - r * 3;
    koff[q] = k < 27 ? c * IPLANE + r * IP + 3 + sx : 0;
  }

  // Ring positions: stem position k (0..5) = stem row y - 1 + k, b0 position m (0..4) = b0 row y - 1 + m.
  // Physical ring rows advance by 4 per group: sb / bb = ring row of position 0.
  int sb = 0, bb = 0;

  // ---- stem rows at positions 2 .. 5: tile = (row pair rp, column block cb); fragment row g = position
  //      2 + 2rp, column 8cb + g; row g + 8 = the position below
  auto stem_phase = [&](const uint16_t* strip, int y) {
    uint2 bf[2][2][3];
    {
      const uint2* s = reinterpret_cast<const uint2*>(sm_raw + OFF_STEM_BF);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 3; ++h) bf[ks][j][h] = s[((ks * 2 + j) * 3 + h) * 32 + lane];
    }
#pragma unroll 1
    for (int cb = warp; cb < CBS; cb += WARPS) {       // both row pairs of the column block: 4 independent MMA chains
      const int js = min(8 * cb + g, NS - 1);
      const bool colv = 8 * cb + g < NS;
      const uint16_t* sbase = strip + 2 * js;
      float acc[2][2][4];
#pragma unroll
      for (int rp = 0; rp < 2; ++rp)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[rp][j][0] = 0.f; acc[rp][j][1] = 0.f; acc[rp][j][2] = 0.f; acc[rp][j][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int rp = 0; rp < 2; ++rp)
#pragma unroll
          for (int q = 0; q < 4; ++q) {                   // a0:(g,k lo) a1:(g+8,k lo) a2:(g,k hi) a3:(g+8,k hi)
            // input row of tap rr for position k = 2 + 2rp + (q & 1): 2*(k - 2) + rr inside the strip
            const uint16_t* sp = sbase + (4 * rp + 2 * (q & 1)) * IP;
            const int ko = ks * 4 + (q >> 1) * 2;
            a[rp][q] = (uint32_t)sp[koff[ko]] | ((uint32_t)sp[koff[ko + 1]] << 16);
          }
        if (ks == 1 && t == 1) {                          // k = 27 carries the bias: A = 1.0
#pragma unroll
          for (int rp = 0; rp < 2; ++rp) {
            a[rp][2] = (a[rp][2] & 0xffffu) | 0x3f800000u;
            a[rp][3] = (a[rp][3] & 0xffffu) | 0x3f800000u;
          }
        }
#pragma unroll
        for (int h = 2; h >= 0; --h)                      // small pieces first
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int rp = 0; rp < 2; ++rp) mma_bf16_16816(acc[rp][j], a[rp], bf[ks][j][h]);
      }
      if (colv) {
        const int scol = 2 * X0 - 2 + js;
        const bool col_in = scol >= 0 && scol < HS;
#pragma unroll
        for (int rp = 0; rp < 2; ++rp)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 + 2 * rp + h;
            const int srow = y - 1 + k;
            const bool inside = col_in && srow >= 0 && srow < HS;
            int slot = sb + k;
            if (slot >= SROWS) slot -= SROWS;
            const uint32_t rowb = sring + (uint32_t)(slot * S_ROWB) + 8 * (t & 1);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float v0 = inside ? fmaxf(acc[rp][j][2 * h], 0.f) : 0.f;
              const float v1 = inside ? fmaxf(acc[rp][j][2 * h + 1], 0.f) : 0.f;
              sts2(rowb + s_off(js, 2 * j + (t >> 1)), v0, v1);
              if (TAPS && inside)
                *reinterpret_cast<float2*>(P.tap_stem + (((size_t)n * HS + srow) * HS + scol) * 16 + 8 * j + 2 * t) =
                    make_float2(v0, v1);
            }
          }
      }
    }
  };

  // ---- b0 rows at positions 1 .. 4 (first = 1: positions 3, 4 only): block = 8 columns x 4 rows, a thread
  //      owns column 8cb + g and channels 4t .. 4t+3 of all four rows
  auto b0_phase = [&](int y, int first) {
    float4 wd[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wd[q] = *reinterpret_cast<const float4*>(&P.dw0[q * 16 + 4 * t]);
    const float4 bd = *reinterpret_cast<const float4*>(&P.dw0[144 + 4 * t]);
    float2 bw[2][2][2];
    {
      const float2* s = reinterpret_cast<const float2*>(sm_raw + OFF_B0_BF);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) bw[ks][j][h] = s[((ks * 2 + j) * 2 + h) * 32 + lane];
    }
    float2 pb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) pb[j] = *reinterpret_cast<const float2*>(sm_raw + OFF_PB0 + (8 * j + 2 * t) * 4);
    uint32_t srow_b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int slot = sb + k;
      if (slot >= SROWS) slot -= SROWS;
      srow_b[k] = sring + (uint32_t)(slot * S_ROWB);
    }
#pragma unroll 1
    for (int cb = warp; cb < CBB; cb += WARPS) {
      const int jb = min(8 * cb + g, NB - 1);
      const bool colv = 8 * cb + g < NB;
      uint32_t coff[3];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) coff[dx] = s_off(jb + dx, t);
      float4 o[4] = {bd, bd, bd, bd};                      // positions 1 .. 4
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (first && k < 2) continue;
        float4 v[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) v[dx] = lds4(srow_b[k] + coff[dx]);
#pragma unroll
        for (int m = 1; m <= 4; ++m) {
          const int r = k - (m - 1);
          if (r < 0 || r > 2) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) fma4(o[m - 1], v[dx], wd[r * 3 + dx]);
        }
      }
      const int bcol = 2 * X0 - 1 + jb;
      const bool col_in = colv && bcol >= 0 && bcol < HS;
      const uint32_t cdst = (uint32_t)((jb & 1) * B_PLANE + (jb >> 1) * 64 + 8 * t);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (first && half == 0) continue;
        float accS[2][4], accB[2][4];
        pw_mma<2>(relu4(o[2 * half]), relu4(o[2 * half + 1]), accS, accB,
                  [&](int ks, int j, int h) { return bw[ks][j][h]; });
        if (colv) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = 1 + 2 * half + h;
            const int brow = y - 1 + m;
            const bool inside = col_in && brow >= 0 && brow < HS;
            int slot = bb + m;
            if (slot >= BROWS) slot -= BROWS;
            const uint32_t dst = bring + (uint32_t)(slot * B_ROWB) + cdst;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float v0 = inside ? fmaxf(accB[j][2 * h] + accS[j][2 * h] + pb[j].x, 0.f) : 0.f;
              const float v1 = inside ? fmaxf(accB[j][2 * h + 1] + accS[j][2 * h + 1] + pb[j].y, 0.f) : 0.f;
              sts2(dst + 32 * j, v0, v1);
              if (TAPS && inside)
                *reinterpret_cast<float2*>(P.tap_b0 + (((size_t)n * HS + brow) * HS + bcol) * 16 + 8 * j + 2 * t) =
                    make_float2(v0, v1);
            }
          }
        }
      }
    }
  };

  // ---- s0.0 output rows y/2 (fragment row g), y/2 + 1 (row g + 8), columns X0 + 8cb + g, from b0 positions 0 .. 4
  auto s0_tile = [&](int y, int bbp, int cb) {
    float4 wd[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wd[q] = *reinterpret_cast<const float4*>(&P.dw1[q * 16 + 4 * t]);
    const float4 bd = *reinterpret_cast<const float4*>(&P.dw1[144 + 4 * t]);
    const float2* sbf = reinterpret_cast<const float2*>(sm_raw + OFF_S0_BF);
    const float* spb = reinterpret_cast<const float*>(sm_raw + OFF_PB1);
    const int pc = 8 * cb + g;
    // column 2*pc + s: s = 0 even plane [pc], 1 odd plane [pc], 2 even plane [pc + 1]
    const uint32_t coff = (uint32_t)(pc * 64 + 16 * t);
    float4 o[2] = {bd, bd};
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      int slot = bbp + m;
      if (slot >= BROWS) slot -= BROWS;
      const uint32_t rowb = bring + (uint32_t)(slot * B_ROWB) + coff;
      const float4 v0 = lds4(rowb), v1 = lds4(rowb + B_PLANE), v2 = lds4(rowb + 64);
#pragma unroll
      for (int yl = 0; yl < 2; ++yl) {
        const int r = m - 2 * yl;
        if (r < 0 || r > 2) continue;
        fma4(o[yl], v0, wd[r * 3 + 0]);
        fma4(o[yl], v1, wd[r * 3 + 1]);
        fma4(o[yl], v2, wd[r * 3 + 2]);
      }
    }
    float accS[5][4], accB[5][4];
    pw_mma<5>(relu4(o[0]), relu4(o[1]), accS, accB,
              [&](int ks, int j, int h) { return sbf[((ks * 5 + j) * 2 + h) * 32 + lane]; });
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* op = P.out + (((size_t)n * HO + (y >> 1) + h) * HO + X0 + pc) * 40 + 2 * t;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float2 pbv = *reinterpret_cast<const float2*>(spb + 8 * j + 2 * t);
        *reinterpret_cast<float2*>(op + 8 * j) =
            make_float2(fmaxf(accB[j][2 * h] + accS[j][2 * h] + pbv.x, 0.f),
                        fmaxf(accB[j][2 * h + 1] + accS[j][2 * h + 1] + pbv.y, 0.f));
      }
    }
  };

#pragma unroll 1
  for (int gi = -1; gi < P.groups; ++gi) {        // gi = -1: the unit's prologue
    const int y = y0 + 4 * gi;
    const int buf = (gi + 1) & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                      // strip(gi) landed; every reader of the previous group is done
    if (gi + 1 < P.groups) fill(buf ^ 1, y + 4);
    stem_phase(strips + buf * STRIP_ELEMS, y);
    __syncthreads();
    b0_phase(y, gi < 0 ? 1 : 0);
    __syncthreads();
    if (gi >= 0 && warp < OW / 8) s0_tile(y, bb, warp);
    sb += 4;
    if (sb >= SROWS) sb -= SROWS;
    bb += 4;
    if (bb >= BROWS) bb -= BROWS;
  }
}

}  // namespace

struct PackedConv {        // device-side packed parameters of one (fused) layer
  float* wpack = nullptr;  // [2][npad_total][kpad]: tf32 hi half, then the residual lo half
  float* bias = nullptr;   // [npad_total]
  float* dw_w = nullptr;   // [9][cin]  (dw-separable only)
  float* dw_b = nullptr;   // [cin]
  CUtensorMap tmW;
  int cin = 0, cout = 0, nt = 0, n_tiles_n = 1, npad_total = 0, mode = LD_PW;
  int stride = 1;          // fixed per layer (decides kc)
  int kc = 32, n_in = 1, taps = 1;
  // input tensor map, re-encoded only when the input pointer / batch changes
  mutable CUtensorMap tmIn, tmInTail;
  mutable const float* tm_in = nullptr;
  mutable int tm_rows = 0, tm_bw = 0, tm_bh = 0;
};

struct DetModel {
  std::map<std::string, PackedConv> conv;
  float* stem_w = nullptr;
  float* stem_b = nullptr;
  float2* b0_bfrag = nullptr;    // b0's 1x1 weights as m16n8k8 tf32 B fragments [kstep][ntile][hi/lo][lane]
  float2* s00_bfrag = nullptr;   // s0.0's 1x1 weights (16 -> 40), same fragment layout with 5 n tiles
  float front_dw[2][160];        // host copies of the b0 / s0.0 depthwise taps [9][16] + bias [16] (kernel parameters)
  const __nv_bfloat16* last_in = nullptr;   // input of the last forward that ran the fused front (det_tap re-runs it)
  int last_n = 0;
  uint2* stem_bfrag = nullptr;   // stem weights + bias row as m16n8k16 B fragments [kstep][ntile][piece][lane]
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
  cudaStream_t aux_stream = nullptr;     // FR_SCRFD_FORK: second branch of the neck / head graph
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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

// w: [cout][cin*ks*ks] (OIHW flattened).  The GEMM K axis is cut into 32-float blocks; block
// kb = (channel block i) * taps + tap holds channels [i*kc, i*kc + kc) of that tap (kc = 16
// leaves the upper half of the block zero; the MMA issuer never reads it).
bool pack(DetModel* m, PackedConv& pc, const std::vector<float>& w, const std::vector<float>& b, int cout, int cin,
          int ks, int mode, int stride) {
  pc.cin = cin;
  pc.cout = cout;
  pc.mode = mode;
  pc.stride = stride;
  pc.kc = (stride == 2 || cin == 16) ? 16 : 32;
  if (mode == LD_PW) pc.kc = 32;
  pc.n_in = (cin + pc.kc - 1) / pc.kc;
  pc.taps = ks * ks;
  const int nkb = pc.n_in * pc.taps;
  const int kpad = nkb * 32;
  pc.npad_total = (cout + 15) / 16 * 16;
  pc.n_tiles_n = pc.npad_total > 256 ? 2 : 1;
  pc.nt = pc.npad_total / pc.n_tiles_n;
  if (pc.nt % 16 != 0 || cin % 4 != 0) return false;
  std::vector<float> wp((size_t)2 * pc.npad_total * kpad, 0.f), bp(pc.npad_total, 0.f);
  for (int co = 0; co < cout; ++co) {
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < pc.taps; ++t) {
        const float v = w[((size_t)co * cin + ci) * pc.taps + t];
        const float h = tf32_rna_host(v);
        const float l = tf32_rna_host(v - h);
        const size_t kk = (size_t)((ci / pc.kc) * pc.taps + t) * 32 + (ci % pc.kc);
        wp[(size_t)co * kpad + kk] = h;
        wp[((size_t)pc.npad_total + co) * kpad + kk] = l;
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
  if (!m->act_allocs.empty()) fr_alloc_epoch()++;
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
  fr_alloc_epoch()++;
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
  if (io.stride != pc.stride) return fr_fail(ctx, FR_ERR_INVALID_ARG, "scrfd layer stride mismatch");
  SepParams p;
  memset(&p, 0, sizeof(p));
  p.out = io.out;
  p.dw_w = pc.dw_w;
  p.dw_b = pc.dw_b;
  p.bias = pc.bias;
  p.add_up = io.add_up;
  p.cin = pc.cin;
  p.cout = pc.cout;
  p.hin = io.hin;
  p.hout = p.wout = io.hin / io.stride;
  p.stride = io.stride;
  p.rows_out = n * p.hout;
  p.mode = pc.mode;
  p.relu = io.relu;
  p.accumulate = io.accumulate;
  p.epi = io.head >= 0 ? EPI_HEAD : EPI_STD;
  if (io.head >= 0) {
    p.score = m->score[io.head];
    p.bbox = m->bbox[io.head];
    p.kps = m->kps[io.head];
  }
  p.kc = pc.kc;
  p.n_in = pc.n_in;
  p.taps = pc.taps;
  p.nt = pc.nt;
  p.n_tiles_n = pc.n_tiles_n;
  p.npad_total = pc.npad_total;
  // output tile: 8 x 16 pixels on the wide maps, 6 x 20 on the 40^2 / 20^2 ones
  if (p.wout % 16 == 0 && p.hout % 8 == 0) { p.th = 8; p.tw = 16; }
  else if (p.wout % 20 == 0) { p.th = 6; p.tw = 20; }
  else return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd feature map size not tileable");
  p.tiles_x = p.wout / p.tw;
  p.tiles_x_magic = p.tiles_x == 1 ? 0u : (uint32_t)((0x100000000ull + p.tiles_x - 1) / p.tiles_x);
  p.hout_magic = (uint32_t)((0x100000000ull + p.hout - 1) / p.hout);
  p.n_shift = pc.n_tiles_n == 2 ? 1 : 0;
  p.num_m_tiles = p.tiles_x * ceil_div(p.rows_out, p.th);
  p.halo = pc.mode == LD_PW ? 0 : 1;
  if (pc.mode == LD_PW) {
    p.bh = p.th;
    p.bw = p.tw;
  } else {
    p.bh = (p.th - 1) * p.stride + 3;
    p.bw = (p.tw - 1) * p.stride + 3;
    // kc = 16, stride 1: a quarter warp reads 4 chunks x 2 adjacent box rows; an odd box width
    // puts those rows on different banks
    if (p.kc == 16 && p.stride == 1 && p.bw % 2 == 0) p.bw++;
  }
  p.in_bytes = (p.bh * p.bw * p.kc * 4 + 1023) / 1024 * 1024;
  p.in_merged = (pc.cin == 16 && pc.kc == 16 && p.bw * 8 <= 256) ? 1 : 0;
  static const bool tail_on = !(getenv("FR_SCRFD_TAIL8") && atoi(getenv("FR_SCRFD_TAIL8")) == 0);   // A/B switch
  p.tail8 = (tail_on && pc.mode == LD_DW && pc.kc == 32 && pc.stride == 1 && pc.n_in > 1 && pc.cin % 32 == 8) ? 1 : 0;
  const int ab_bytes = 2 * A_BYTES + 2 * pc.nt * 128;
  const int dw_bytes = pc.mode == LD_DW ? (10 * pc.cin * 4 + 15) / 16 * 16 : 0;
  const int budget = 227 * 1024 - 1024 - 512 - dw_bytes - pc.npad_total * 4;
  p.s_in = 2;
  p.s_ab = 2;
  if (p.s_in * p.in_bytes + p.s_ab * ab_bytes > budget)
    return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd layer does not fit shared memory");
  // A/B stages to 3, then input boxes (TMA latency under load is ~3 us: keep many in flight), then A/B to 4
  auto fits = [&](int add) { return p.s_in * p.in_bytes + p.s_ab * ab_bytes + add <= budget; };
  if (fits(ab_bytes)) p.s_ab++;
  while (p.s_in < MAX_STAGES && fits(p.in_bytes)) p.s_in++;
  if (p.s_ab < 4 && fits(ab_bytes)) p.s_ab++;
  // a second hi*hi accumulator only where one would take more than 12 truncating accumulations
  p.nbig = pc.cin * pc.taps <= 96 ? 1 : 2;
  const int slot = (p.nbig + 1) * pc.nt;
  if (slot > 512) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd layer does not fit tensor memory");
  p.acc_stages = 2 * slot <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * slot) cols *= 2;
  p.tmem_cols = cols;
  p.err_flag = m->err_flag;
  const int rows_in = n * io.hin;
  if (pc.tm_in != io.in || pc.tm_rows != rows_in || pc.tm_bw != p.bw || pc.tm_bh != p.bh) {
    const bool ok_map = p.in_merged
        ? tc_make_map_2d_u64(&pc.tmIn, io.in, (uint64_t)io.hin * 8, (uint64_t)rows_in, (uint32_t)p.bw * 8, (uint32_t)p.bh)
        : tc_make_map_3d_f32(&pc.tmIn, io.in, (uint64_t)pc.cin, (uint64_t)io.hin, (uint64_t)rows_in, (uint32_t)p.kc,
                             (uint32_t)p.bw, (uint32_t)p.bh);
    if (!ok_map)
      return fr_fail(ctx, FR_ERR_CUDA, "scrfd input tensor map creation failed");
    pc.tmInTail = pc.tmIn;
    if (p.tail8 && !tc_make_map_3d_f32(&pc.tmInTail, io.in, (uint64_t)pc.cin, (uint64_t)io.hin, (uint64_t)rows_in, 8u,
                                       (uint32_t)p.bw, (uint32_t)p.bh))
      return fr_fail(ctx, FR_ERR_CUDA, "scrfd input tensor map creation failed");
    pc.tm_in = io.in;
    pc.tm_rows = rows_in;
    pc.tm_bw = p.bw;
    pc.tm_bh = p.bh;
  }
  const int smem = p.s_in * p.in_bytes + p.s_ab * ab_bytes + 1024 + 512 + dw_bytes + pc.npad_total * 4;
  const int total_tiles = p.num_m_tiles * p.n_tiles_n;
  const int grid = std::min(total_tiles, m->num_sms);
  static const char* dbg_env = getenv("FR_SCRFD_DBG");   // "<launch index>:<file>"
  static int dbg_count = 0;
  long long* dbg_buf = nullptr;
  if (dbg_env && atoi(dbg_env) == dbg_count++) {
    cudaMallocManaged(&dbg_buf, 5 * 256 * 2 * sizeof(long long));
    memset(dbg_buf, 0, 5 * 256 * 2 * sizeof(long long));
    p.dbg = dbg_buf;
  }
  sep_gemm_kernel<<<grid, THREADS, smem, ctx->stream>>>(pc.tmIn, pc.tmInTail, pc.tmW, p);
  if (dbg_buf) {
    cudaStreamSynchronize(ctx->stream);
    FILE* f = fopen(strchr(dbg_env, ':') + 1, "w");
    if (f) {
      fprintf(f, "# nkb_per_tile=%d n_in=%d taps=%d nt=%d s_in=%d s_ab=%d kc=%d\n", pc.n_in * pc.taps, pc.n_in, pc.taps, pc.nt, p.s_in, p.s_ab, p.kc);
      for (int r = 0; r < 5; ++r)
        for (int i = 0; i < 256; ++i) fprintf(f, "%d %d %lld %lld\n", r, i, dbg_buf[(r * 256 + i) * 2], dbg_buf[(r * 256 + i) * 2 + 1]);
      fclose(f);
    }
    cudaFree(dbg_buf);
  }
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

// dense 3x3 stride-1 conv through conv3_halo_kernel (head >= 0: the fused 64 -> 30 head conv)
int launch_conv3_halo(fr_ctx* ctx, const PackedConv& pc, const float* in, float* out, int hw, int head, int n) {
  DetModel* m = ctx->det;
  if (pc.mode != LD_IM2COL || pc.stride != 1 || pc.n_tiles_n != 1)
    return fr_fail(ctx, FR_ERR_INVALID_ARG, "conv3_halo: not a dense 3x3 stride-1 layer");
  C3Params p;
  memset(&p, 0, sizeof(p));
  p.out = out;
  p.bias = pc.bias;
  p.epi = head >= 0 ? EPI_HEAD : EPI_STD;
  if (head >= 0) {
    p.score = m->score[head];
    p.bbox = m->bbox[head];
    p.kps = m->kps[head];
  }
  p.cin = pc.cin;
  p.cout = pc.cout;
  p.h = p.w = hw;
  p.n_img = n;
  p.kc = pc.kc;
  p.n_in = pc.n_in;
  p.nt = pc.nt;
  p.npad_total = pc.npad_total;
  if (hw % 16 == 0 && hw >= 80) { p.tw = 16; p.th = 7; }
  else if (hw % 20 == 0) { p.tw = 20; p.th = 5; }
  else return fr_fail(ctx, FR_ERR_UNSUPPORTED, "scrfd feature map size not tileable");
  p.tw2 = p.tw + 2;
  p.tiles_x = hw / p.tw;
  p.tiles_y = ceil_div(hw, p.th);
  p.num_tiles = n * p.tiles_x * p.tiles_y;
  auto magic = [](uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d); };
  p.tiles_per_img_magic = magic((uint32_t)(p.tiles_x * p.tiles_y));
  p.tiles_x_magic = magic((uint32_t)p.tiles_x);
  p.tw2_magic = magic((uint32_t)p.tw2);
  p.chunk_magic = magic((uint32_t)(p.kc / 4));
  p.box_px = (p.th + 2) * p.tw2;
  p.in_bytes = (p.box_px * p.kc * 4 + 1023) / 1024 * 1024;
  p.a_rows = (127 + 2 * p.tw2 + 2 + 1 + 7) / 8 * 8;
  const int a2 = 2 * p.a_rows * 128, b2 = 2 * pc.nt * 128, nkb = pc.n_in * 9;
  const int budget = 227 * 1024 - 1024 - 1024 - pc.npad_total * 4;
  p.s_in = 2;
  p.s_a = 2;
  const int rest = budget - p.s_in * p.in_bytes - p.s_a * a2;
  p.b_resident = (nkb <= C3_MAX_B && nkb * b2 <= rest) ? 1 : 0;
  p.s_b = p.b_resident ? nkb : std::min(10, rest / b2);   // streamed weights: keep ~one box of taps in flight
  if (p.s_b < 2) return fr_fail(ctx, FR_ERR_UNSUPPORTED, "conv3_halo: does not fit shared memory");
  auto used = [&]() { return p.s_in * p.in_bytes + p.s_a * a2 + p.s_b * b2; };
  if (used() + p.in_bytes <= budget) p.s_in++;
  if (used() + a2 <= budget) p.s_a++;
  if (used() + p.in_bytes <= budget) p.s_in++;
  const int slot = 5 * pc.nt;   // hh0 | hl0 | hh1 | hl1 | lh
  p.acc_stages = 2 * slot <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * slot) cols *= 2;
  p.tmem_cols = cols;
  p.err_flag = m->err_flag;
  const int rows_in = n * hw;
  if (pc.tm_in != in || pc.tm_rows != rows_in || pc.tm_bw != p.tw2 || pc.tm_bh != p.th + 2) {
    if (!tc_make_map_3d_f32(&pc.tmIn, in, (uint64_t)pc.cin, (uint64_t)hw, (uint64_t)rows_in, (uint32_t)p.kc,
                            (uint32_t)p.tw2, (uint32_t)(p.th + 2)))
      return fr_fail(ctx, FR_ERR_CUDA, "scrfd input tensor map creation failed");
    pc.tm_in = in;
    pc.tm_rows = rows_in;
    pc.tm_bw = p.tw2;
    pc.tm_bh = p.th + 2;
  }
  const int smem = used() + 1024 + 1024 + pc.npad_total * 4;
  const int grid = std::min(p.num_tiles, m->num_sms);
  conv3_halo_kernel<<<grid, THREADS, smem, ctx->stream>>>(pc.tmIn, pc.tmW, p);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

}  // namespace

int det_model_create(fr_ctx* ctx, const fr_weights* w) {
  if (!w || w->model != FR_MODEL_DET) return fr_fail(ctx, FR_ERR_MODEL, "det weights missing");
  std::unique_ptr<DetModel> m(new DetModel());
  bool ok = true;
  auto dense = [&](const std::string& name, int stride) {   // conv (3x3 or 1x1) as a [cout][cin*k*k] GEMM
    const fr_tensor& tw = w->at(name + ".w");
    const int cout = (int)tw.dims[0], cin = (int)tw.dims[1], k = (int)tw.dims[2];
    ok = ok && pack(m.get(), m->conv[name], tw.data, w->at(name + ".b").data, cout, cin, k,
                    k == 3 ? LD_IM2COL : LD_PW, stride);
  };
  auto dwsep = [&](const std::string& name, int stride) {   // dw 3x3 + ReLU built inside the 1x1's operand
    const fr_tensor& pw = w->at(name + ".pw.w");
    const int cout = (int)pw.dims[0], cin = (int)pw.dims[1];
    PackedConv& pc = m->conv[name];
    ok = ok && pack(m.get(), pc, pw.data, w->at(name + ".pw.b").data, cout, cin, 1, LD_DW, stride);
    const std::vector<float>& dw = w->at(name + ".dw.w").data;   // [cin][1][3][3] -> [9][cin]
    std::vector<float> dwt((size_t)9 * cin);
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) dwt[(size_t)t * cin + c] = dw[(size_t)c * 9 + t];
    pc.dw_w = upload(m.get(), dwt);
    pc.dw_b = upload(m.get(), w->at(name + ".dw.b").data);
    ok = ok && pc.dw_w && pc.dw_b;
    if ((name == "b0" || name == "s0.0") && cin == 16) {   // kernel-parameter copies for front_fused_kernel
      float* dst = m->front_dw[name == "b0" ? 0 : 1];
      memcpy(dst, dwt.data(), 144 * sizeof(float));
      memcpy(dst + 144, w->at(name + ".dw.b").data.data(), 16 * sizeof(float));
    }
  };
  {
    const fr_tensor& tw = w->at("stem.w");   // [16][3][3][3] -> [27][16]
    std::vector<float> sw(27 * 16);
    for (int co = 0; co < 16; ++co)
      for (int k = 0; k < 27; ++k) sw[k * 16 + co] = tw.data[co * 27 + k];
    m->stem_w = upload(m.get(), sw);
    m->stem_b = upload(m.get(), w->at("stem.b").data);
    ok = ok && m->stem_w && m->stem_b;
    // B fragments (lane = 4g + t): b0 = W[k0 + 2t, +1][n = 8j + g], b1 = W[k0 + 2t + 8, +9][n]; row 27 = bias;
    // each fp32 value as three bf16 pieces (hi, mid, lo)
    const std::vector<float>& sbv = w->at("stem.b").data;
    auto piece = [&](int k, int nn, int h) -> uint32_t {
      float v = k < 27 ? sw[(size_t)k * 16 + nn] : (k == 27 ? sbv[nn] : 0.f);
      __nv_bfloat16 pb = __float2bfloat16_rn(v);
      for (int i = 0; i < h; ++i) {
        v -= __bfloat162float(pb);
        pb = __float2bfloat16_rn(v);
      }
      uint16_t u;
      memcpy(&u, &pb, 2);
      return u;
    };
    std::vector<uint2> frag(2 * 2 * 3 * 32);
    for (int ks = 0; ks < 2; ++ks)
      for (int j = 0; j < 2; ++j)
        for (int h = 0; h < 3; ++h)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, k0 = ks * 16 + 2 * t, nn = 8 * j + g;
            uint2 f;
            f.x = piece(k0, nn, h) | (piece(k0 + 1, nn, h) << 16);
            f.y = piece(k0 + 8, nn, h) | (piece(k0 + 9, nn, h) << 16);
            frag[((ks * 2 + j) * 3 + h) * 32 + lane] = f;
          }
    if (cudaMalloc(&m->stem_bfrag, frag.size() * sizeof(uint2)) != cudaSuccess) ok = false;
    else {
      cudaMemcpy(m->stem_bfrag, frag.data(), frag.size() * sizeof(uint2), cudaMemcpyHostToDevice);
      m->allocs.push_back(m->stem_bfrag);
    }
  }
  dwsep("b0", 1);
  {
    // b0's 1x1 for b0_dwpw_mma_kernel: B fragment of lane 4g + t = W[n = 8j + g][channel 4t + 2ks (+1)], hi / lo tf32
    const fr_tensor& pw = w->at("b0.pw.w");
    ok = ok && pw.dims[0] == 16 && pw.dims[1] == 16;
    std::vector<float2> frag(2 * 2 * 2 * 32);
    for (int ks = 0; ok && ks < 2; ++ks)
      for (int j = 0; j < 2; ++j)
        for (int h = 0; h < 2; ++h)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, nn = 8 * j + g;
            float v[2];
            for (int i = 0; i < 2; ++i) {
              const float wv = pw.data[(size_t)nn * 16 + 4 * t + 2 * ks + i];
              const float hi = tf32_rna_host(wv);
              v[i] = h == 0 ? hi : tf32_rna_host(wv - hi);
            }
            frag[((ks * 2 + j) * 2 + h) * 32 + lane] = make_float2(v[0], v[1]);
          }
    if (ok && cudaMalloc(&m->b0_bfrag, frag.size() * sizeof(float2)) == cudaSuccess) {
      cudaMemcpy(m->b0_bfrag, frag.data(), frag.size() * sizeof(float2), cudaMemcpyHostToDevice);
      m->allocs.push_back(m->b0_bfrag);
    } else {
      ok = false;
    }
  }
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kStages[s][0]; ++b) dwsep("s" + std::to_string(s) + "." + std::to_string(b), b == 0 ? 2 : 1);
  {
    // s0.0's 1x1 for front_fused_kernel: B fragment of lane 4g + t = W[n = 8j + g][channel 4t + 2ks (+1)], hi / lo tf32
    const fr_tensor& pw = w->at("s0.0.pw.w");
    ok = ok && pw.dims[0] == 40 && pw.dims[1] == 16;
    std::vector<float2> frag(2 * 5 * 2 * 32);
    for (int ks = 0; ok && ks < 2; ++ks)
      for (int j = 0; j < 5; ++j)
        for (int h = 0; h < 2; ++h)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, nn = 8 * j + g;
            float v[2];
            for (int i = 0; i < 2; ++i) {
              const float wv = pw.data[(size_t)nn * 16 + 4 * t + 2 * ks + i];
              const float hi = tf32_rna_host(wv);
              v[i] = h == 0 ? hi : tf32_rna_host(wv - hi);
            }
            frag[((ks * 5 + j) * 2 + h) * 32 + lane] = make_float2(v[0], v[1]);
          }
    if (ok && cudaMalloc(&m->s00_bfrag, frag.size() * sizeof(float2)) == cudaSuccess) {
      cudaMemcpy(m->s00_bfrag, frag.data(), frag.size() * sizeof(float2), cudaMemcpyHostToDevice);
      m->allocs.push_back(m->s00_bfrag);
    } else {
      ok = false;
    }
  }
  for (int i = 0; i < 3; ++i) dense("lat" + std::to_string(i), 1);
  for (int i = 0; i < 3; ++i) dense("fpn" + std::to_string(i), 1);
  for (int i = 0; i < 2; ++i) dense("down" + std::to_string(i), 2);
  for (int i = 0; i < 2; ++i) dense("pafpn" + std::to_string(i), 1);
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    dwsep(h + ".t0", 1);
    dwsep(h + ".t1", 1);
    // fuse the three head convs of the stride into one 64 -> 30 conv (cls 2 | reg 8 | kps 20)
    std::vector<float> fw, fb;
    for (const char* part : {".cls", ".reg", ".kps"}) {
      const fr_tensor& tw = w->at(h + part + ".w");
      const fr_tensor& tb = w->at(h + part + ".b");
      fw.insert(fw.end(), tw.data.begin(), tw.data.end());
      fb.insert(fb.end(), tb.data.begin(), tb.data.end());
    }
    ok = ok && pack(m.get(), m->conv[h + ".out"], fw, fb, 30, 64, 3, LD_IM2COL, 1);
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
    ok = cudaFuncSetAttribute(sep_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
         cudaFuncSetAttribute(conv3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
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
  if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  delete m;
  ctx->det = nullptr;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// The 320^2 / 160^2 activations are 6.5 / 4.1 MB per frame in fp32: at 64 frames every early layer
// streams 260-840 MB through HBM.  The first FR_DET_CHUNK_LAYERS layers (stem, b0, s0.0, s0.1, ...)
// therefore run depth-first over chunks of FR_DET_CHUNK frames whose tensors stay in the 126 MB L2:
// each chunk uses frames [0, chunk) of the activation buffers and the last chunked layer writes its
// output at the chunk's offset of the full-batch tensor.  Per-frame results do not depend on the
// batch around a frame, so this is bit-identical to the unchunked run.
// stem -> b0 -> s0.0 in one launch (front_fused_kernel); taps: also write the stem / b0 activations
static int launch_front_fused(fr_ctx* ctx, const __nv_bfloat16* d_in_chw, int n, bool taps) {
  DetModel* m = ctx->det;
  const PackedConv& b0 = m->conv.at("b0");
  const PackedConv& s00 = m->conv.at("s0.0");
  ff::Params P;
  P.in = d_in_chw;
  P.out = m->a_stage[0];
  P.stem_bf = m->stem_bfrag;
  P.b0_bf = m->b0_bfrag;
  P.s0_bf = m->s00_bfrag;
  P.pb0 = b0.bias;
  P.pb1 = s00.bias;
  memcpy(P.dw0, m->front_dw[0], sizeof(P.dw0));
  memcpy(P.dw1, m->front_dw[1], sizeof(P.dw1));
  P.tap_stem = taps ? m->a_stem : nullptr;
  P.tap_b0 = taps ? m->a_b0 : nullptr;
  // groups per unit: long units amortise the prologue (3 stem rows + 1 b0 row per unit); small batches
  // get short units so that one frame still spreads over the whole GPU
  P.groups = n >= 8 ? 10 : 2;
  const int units = n * 4 * ((ff::HS / 4) / P.groups);
  if (taps) {
    FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, front_fused_kernel<true>, ff::SMEM_BYTES));
    front_fused_kernel<true><<<(unsigned)units, ff::FTHREADS, ff::SMEM_BYTES, ctx->stream>>>(P);
  } else {
    FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, front_fused_kernel<false>, ff::SMEM_BYTES));
    front_fused_kernel<false><<<(unsigned)units, ff::FTHREADS, ff::SMEM_BYTES, ctx->stream>>>(P);
  }
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

static bool front_fused_enabled() {
  const char* v = getenv("FR_SCRFD_FRONT_FUSED");   // A/B switch: 0 = the three separate kernels
  return !(v && atoi(v) == 0);
}

static int det_front(fr_ctx* ctx, const __nv_bfloat16* d_in_chw, int n, int n_layers, size_t out_frame_off) {
  DetModel* m = ctx->det;
  m->last_in = nullptr;
  if (n_layers >= 15 && out_frame_off == 0 && front_fused_enabled()) {
    FR_CHECK(launch_front_fused(ctx, d_in_chw, n, false));
    m->last_in = d_in_chw;
    m->last_n = n;
    const float* cur = m->a_stage[0];
    int hw = 160, bi = 1;
    for (int s = 0; s < 4; ++s)
      for (int b = (s == 0 ? 1 : 0); b < kStages[s][0]; ++b, ++bi) {
        const int stride = b == 0 ? 2 : 1;
        const int ho = hw / stride;
        LayerIO io;
        io.in = cur; io.out = m->a_stage[bi]; io.hin = hw; io.stride = stride; io.relu = 1;
        FR_CHECK(launch_layer(ctx, m->conv.at("s" + std::to_string(s) + "." + std::to_string(b)), io, n));
        hw = ho;
        cur = m->a_stage[bi];
      }
    return FR_OK;
  }
  // last front layer writes at frame offset `out_frame_off` of its (full-batch) output tensor
  auto dst = [&](int layer, float* base, size_t per_frame) {
    return layer == n_layers - 1 ? base + out_frame_off * per_frame : base;
  };
  // stem: 3x3 s2, 3 -> 16, ReLU (bf16 planar input from K1) -> fp32 NHWC
  float* o_stem = dst(0, m->a_stem, (size_t)16 * 320 * 320);
  {
    static const bool simt_stem = getenv("FR_SCRFD_STEM_SIMT") != nullptr;   // A/B switch: the CUDA-core stem
    if (simt_stem) {
      const size_t total = (size_t)n * (DET / 2) * (DET / 2 / 4);
      stem_conv_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(d_in_chw, o_stem, m->stem_w,
                                                                                  m->stem_b, n);
    } else {
      FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, stem_conv_mma_kernel, SM_SMEM_BYTES));
      stem_conv_mma_kernel<<<(unsigned)(n * (DET / (4 * SM_GROUPS))), 128, SM_SMEM_BYTES, ctx->stream>>>(
          d_in_chw, o_stem, m->stem_bfrag);
    }
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
  }
  if (n_layers == 1) return FR_OK;
  auto dwsep = [&](const std::string& name, const float* in, float* out, int hin, int stride) -> int {
    LayerIO io;
    io.in = in; io.out = out; io.hin = hin; io.stride = stride; io.relu = 1;
    return launch_layer(ctx, m->conv.at(name), io, n);
  };
  float* o_b0 = dst(1, m->a_b0, (size_t)16 * 320 * 320);
  static const bool b0_tc = getenv("FR_SCRFD_B0_TCGEN05") != nullptr;   // A/B switch: b0 through sep_gemm_kernel
  if (b0_tc) {
    FR_CHECK(dwsep("b0", m->a_stem, o_b0, 320, 1));
  } else {
    const PackedConv& pc = m->conv.at("b0");
    FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, b0_dwpw_mma_kernel, B0_SMEM_BYTES));
    constexpr int HW = DET / 2;
    b0_dwpw_mma_kernel<<<(unsigned)(n * (HW / B0_TW) * (HW / (B0_ROWS * B0_GROUPS))), 256, B0_SMEM_BYTES, ctx->stream>>>(
        m->a_stem, o_b0, pc.dw_w, pc.dw_b, m->b0_bfrag, pc.bias, HW);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
  }
  const float* cur = m->a_b0;
  int hw = 320, bi = 0;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kStages[s][0]; ++b, ++bi) {
      if (bi + 2 >= n_layers) return FR_OK;
      const int stride = b == 0 ? 2 : 1;
      const int ho = hw / stride;
      FR_CHECK(dwsep("s" + std::to_string(s) + "." + std::to_string(b), cur,
                     dst(bi + 2, m->a_stage[bi], (size_t)kStages[s][1] * ho * ho), hw, stride));
      hw = ho;
      cur = m->a_stage[bi];
    }
  return FR_OK;
}

int det_forward(fr_ctx* ctx, const __nv_bfloat16* d_in_chw, int n, HeadPtrs* heads) {
  DetModel* m = ctx->det;
  if (!m) return fr_fail(ctx, FR_ERR_NOT_LOADED, "Model not loaded!");
  if (n <= 0) return FR_OK;
  int cap = 1;
  while (cap < n) cap *= 2;
  FR_CHECK(det_build_acts(ctx, cap));
  static const int chunk = env_int("FR_DET_CHUNK", 0);
  static const int chunk_layers = std::min(std::max(env_int("FR_DET_CHUNK_LAYERS", 4), 1), 15);
  const int n_backbone = 15;   // stem, b0, 13 dw-separable blocks
  int front = n_backbone;      // layers run by det_front
  if (chunk > 0 && n > chunk) {
    for (int f0 = 0; f0 < n; f0 += chunk)
      FR_CHECK(det_front(ctx, d_in_chw + (size_t)f0 * 3 * DET * DET, std::min(chunk, n - f0), chunk_layers, (size_t)f0));
    front = chunk_layers;
  } else {
    FR_CHECK(det_front(ctx, d_in_chw, n, n_backbone, 0));
  }
  auto dwsep = [&](const std::string& name, const float* in, float* out, int hin, int stride) -> int {
    LayerIO io;
    io.in = in; io.out = out; io.hin = hin; io.stride = stride; io.relu = 1;
    return launch_layer(ctx, m->conv.at(name), io, n);
  };
  // the rest of the backbone on the whole batch
  const float* cur = m->a_b0;
  int hw = 320, bi = 0;
  const float* feats[3] = {nullptr, nullptr, nullptr};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < kStages[s][0]; ++b, ++bi) {
      const int stride = b == 0 ? 2 : 1;
      if (bi + 2 >= front)
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
    static const bool use_halo = !getenv("FR_SCRFD_NO_C3HALO");
    if (stride == 1 && !accumulate && use_halo) return launch_conv3_halo(ctx, m->conv.at(name), in, out, hin, -1, n);
    LayerIO io;
    io.in = in; io.out = out; io.hin = hin; io.stride = stride; io.accumulate = accumulate;
    return launch_layer(ctx, m->conv.at(name), io, n);
  };
  const float* outs[3] = {m->inter[0], m->pout[1], m->pout[2]};
  auto head = [&](int i) -> int {
    const std::string h = "h" + std::to_string(i);
    FR_CHECK(dwsep(h + ".t0", outs[i], m->tw0[i], fh[i], 1));
    FR_CHECK(dwsep(h + ".t1", m->tw0[i], m->tw1[i], fh[i], 1));
    static const bool use_halo = !getenv("FR_SCRFD_NO_C3HALO");
    if (use_halo) {
      FR_CHECK(launch_conv3_halo(ctx, m->conv.at(h + ".out"), m->tw1[i], nullptr, fh[i], i, n));
    } else {
      LayerIO io;
      io.in = m->tw1[i]; io.hin = fh[i]; io.head = i;
      FR_CHECK(launch_layer(ctx, m->conv.at(h + ".out"), io, n));
    }
    heads->score[i] = m->score[i];
    heads->bbox[i] = m->bbox[i];
    heads->kps[i] = m->kps[i];
    return FR_OK;
  };
  FR_CHECK(conv3("fpn0", m->lat[0], m->inter[0], fh[0], 1, 0));
  // FR_SCRFD_FORK=1 (experiment): the stride-8 head (3 launches, ~325 us) depends on fpn0 only; on a second stream
  // its CTAs fill the SMs that the short 40^2 / 20^2 launches of the other branch leave idle in their last wave
  static const bool fork = getenv("FR_SCRFD_FORK") && atoi(getenv("FR_SCRFD_FORK")) != 0;
  cudaStream_t main_stream = ctx->stream;
  if (fork) {
    if (!m->aux_stream) {
      FR_CUDA_OK(ctx, cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
      FR_CUDA_OK(ctx, cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
      FR_CUDA_OK(ctx, cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
    FR_CUDA_OK(ctx, cudaEventRecord(m->ev_fork, main_stream));
    FR_CUDA_OK(ctx, cudaStreamWaitEvent(m->aux_stream, m->ev_fork, 0));
    ctx->stream = m->aux_stream;
    const int s = head(0);
    ctx->stream = main_stream;
    FR_CHECK(s);
    FR_CUDA_OK(ctx, cudaEventRecord(m->ev_join, m->aux_stream));
  }
  for (int i = 1; i < 3; ++i) FR_CHECK(conv3("fpn" + std::to_string(i), m->lat[i], m->inter[i], fh[i], 1, 0));
  for (int i = 0; i < 2; ++i)
    FR_CHECK(conv3("down" + std::to_string(i), m->inter[i], m->inter[i + 1], fh[i], 2, 1));
  for (int i = 1; i < 3; ++i)
    FR_CHECK(conv3("pafpn" + std::to_string(i - 1), m->inter[i], m->pout[i], fh[i], 1, 0));
  for (int i = fork ? 1 : 0; i < 3; ++i) FR_CHECK(head(i));
  if (fork) FR_CUDA_OK(ctx, cudaStreamWaitEvent(main_stream, m->ev_join, 0));
  return FR_OK;
}

// Test hook: copy one intermediate activation (fp32 NHWC) of the last det_forward to the host.
// tap: 0 stem, 1 b0, 2..14 backbone blocks, 15..17 lat, 18..20 inter, 21..22 pout[1..2],
// 23..25 head tower 0, 26..28 head tower 1.
int det_tap(fr_ctx* ctx, int tap, int n, float* h_out, size_t out_elems) {
  DetModel* m = ctx->det;
  if (!m || m->cap < n) return fr_fail(ctx, FR_ERR_NOT_LOADED, "no detector activations");
  const float* src = nullptr;
  // the fused front keeps the stem / b0 activations in shared memory: re-run it with the tap copies on
  if (tap <= 1 && m->last_in && m->last_n >= n) FR_CHECK(launch_front_fused(ctx, m->last_in, m->last_n, true));
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
