// K3 scrfd_decode + K4 nms.
//
// K3: three-stride anchor decode (north_star; InsightFace distance2bbox/distance2kps
// semantics) producing exactly the rows FaceDetector::postprocess consumes
// (reference src/face_detector.cpp:242-278): score > thr (strict, :253), x/scale IEEE divide
// (:255-258), cv::Rect(int(x1), int(y1), int(x2-x1), int(y2-y1)) truncation (:260-265),
// landmarks/scale (:270-273).  Only the score planes are scanned (67,200 B/frame); survivors
// are gathered.
// K4: FaceDetector::nms / iou (src/face_detector.cpp:340-384): sort by score descending
// (canonical tie order: anchor index ascending), integer IoU, strict '>' suppression, output
// in sorted order.  One CTA per frame; 32-candidate blocks are resolved by one warp with
// ballots/shuffles, then the surviving boxes of the block suppress the rest in parallel.
#include "common.h"

namespace {

constexpr int NA = FR_NUM_ANCHORS;     // 16800
constexpr int KEY_STRIDE = 32768;      // per-image key capacity (power of two >= NA)
constexpr int SMEM_CAND = 4096;        // candidates handled entirely in shared memory
constexpr int MAT_CAND = 512;          // ... and with the bit-matrix NMS (512 x 16 words behind the first 512 boxes)
static_assert(MAT_CAND * 16 + MAT_CAND * (MAT_CAND / 32) * 4 <= SMEM_CAND * 16, "bit matrix aliases the unused boxes");
constexpr int NMS_THREADS = 1024;

struct DecodeArgs {
  HeadPtrs h;
  const ImgDesc* desc;
  const float* scales;
};

__device__ __forceinline__ void anchor_locate(int a, int& s, int& local, int& n_s, int& ws,
                                              int& stride) {
  if (a < 12800) { s = 0; local = a; n_s = 12800; ws = 80; stride = 8; }
  else if (a < 16000) { s = 1; local = a - 12800; n_s = 3200; ws = 40; stride = 16; }
  else { s = 2; local = a - 16000; n_s = 800; ws = 20; stride = 32; }
}

__device__ __forceinline__ unsigned int float_sortable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unsortable(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void decode_threshold_kernel(DecodeArgs args, float thr, unsigned long long* keys,
                                        int* counts) {
  const int img = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  float score = 0.f;
  if (a < NA) {
    int s, local, n_s, ws, stride;
    anchor_locate(a, s, local, n_s, ws, stride);
    score = __ldg(args.h.score[s] + (size_t)img * n_s + local);
    hit = score > thr;  // strict, face_detector.cpp:253
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(&counts[img], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (hit) {
    const int pos = base + __popc(m & ((1u << lane) - 1));
    keys[(size_t)img * KEY_STRIDE + pos] =
        ((unsigned long long)float_sortable(score) << 32) | (unsigned int)(~(unsigned int)a);
  }
}

__device__ __forceinline__ int4 decode_box(const DecodeArgs& args, int img, int a, float scale) {
  int s, local, n_s, ws, stride;
  anchor_locate(a, s, local, n_s, ws, stride);
  const int cell = local >> 1;
  const float cx = (float)((cell % ws) * stride);
  const float cy = (float)((cell / ws) * stride);
  const float st = (float)stride;
  const float4 d = __ldg(reinterpret_cast<const float4*>(args.h.bbox[s]) + (size_t)img * n_s + local);
  const float x1 = __fdiv_rn(__fsub_rn(cx, __fmul_rn(d.x, st)), scale);
  const float y1 = __fdiv_rn(__fsub_rn(cy, __fmul_rn(d.y, st)), scale);
  const float x2 = __fdiv_rn(__fadd_rn(cx, __fmul_rn(d.z, st)), scale);
  const float y2 = __fdiv_rn(__fadd_rn(cy, __fmul_rn(d.w, st)), scale);
  int4 r;
  r.x = __float2int_rz(x1);
  r.y = __float2int_rz(y1);
  r.z = __float2int_rz(__fsub_rn(x2, x1));
  r.w = __float2int_rz(__fsub_rn(y2, y1));
  return r;
}

__device__ __forceinline__ bool iou_gt(const int4& a, const int4& b, float thr) {
  const int x1 = max(a.x, b.x), y1 = max(a.y, b.y);
  const int x2 = min(a.x + a.z, b.x + b.z), y2 = min(a.y + a.w, b.y + b.w);
  const int w = max(0, x2 - x1), h = max(0, y2 - y1);
  const int inter = w * h;
  const int den = a.z * a.w + b.z * b.w - inter;
  return __fdiv_rn((float)inter, (float)den) > thr;  // 0/0 -> NaN -> false
}

// dynamic smem: keys[SMEM_CAND] (u64) | boxes[SMEM_CAND] (int4) | supp[NA] (u8) | keep[] | pref[]
__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(DecodeArgs args, float nms_thr, unsigned long long* keys_g, int4* boxes_g,
           const int* counts, fr_face* out, int cap, int* n_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* keys_s = reinterpret_cast<unsigned long long*>(smem);
  int4* boxes_s = reinterpret_cast<int4*>(smem + SMEM_CAND * 8);
  unsigned char* supp = smem + SMEM_CAND * 8 + SMEM_CAND * 16;
  unsigned int* keep = reinterpret_cast<unsigned int*>(supp + ((NA + 15) / 16) * 16);
  int* pref = reinterpret_cast<int*>(keep + (NA / 32 + 2));
  __shared__ unsigned int blockmask;
  __shared__ int kept_total;
  __shared__ int cur_block;

  const int img = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = min(counts[img], NA);
  const float scale = args.scales ? args.scales[img] : args.desc[img].scale;
  if (n == 0) {
    if (tid == 0) n_out[img] = 0;
    return;
  }
  const bool in_smem = n <= SMEM_CAND;
  unsigned long long* K = in_smem ? keys_s : keys_g + (size_t)img * KEY_STRIDE;
  int4* B = in_smem ? boxes_s : boxes_g + (size_t)img * NA;
  int P = 1;
  while (P < n) P <<= 1;
  // load + pad
  for (int i = tid; i < P; i += NMS_THREADS) {
    unsigned long long k = i < n ? keys_g[(size_t)img * KEY_STRIDE + i] : 0ull;
    K[i] = k;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += NMS_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = K[i], b = K[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { K[i] = b; K[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += NMS_THREADS) {
    const int a = (int)(~(unsigned int)(K[i] & 0xffffffffull));
    B[i] = decode_box(args, img, a, scale);
    supp[i] = 0;
  }
  __syncthreads();

  const int nblk = (n + 31) >> 5;
  if (n <= MAT_CAND) {
    // Few candidates (the usual case): the whole suppression relation as a bit matrix, in parallel
    // (word (i, w) = candidates j of word w behind i with IoU(i, j) > thr: independent IoU evaluations, no
    // serial chain), then ONE warp walks the candidates in score order with the 'removed' set in registers
    // (lane w = word w): candidate i is kept iff no earlier kept candidate removed it -- the same greedy
    // rule, ~30 cycles per candidate instead of a ballot / shuffle / IoU chain per keeper.
    unsigned int* mat = reinterpret_cast<unsigned int*>(boxes_s + MAT_CAND);      // [n][nblk], behind the n boxes
    for (int e = tid; e < n * nblk; e += NMS_THREADS) {
      const int i = e / nblk, w = e - i * nblk;
      unsigned int bits = 0;
      if ((w << 5) + 31 > i) {
        const int4 bi = B[i];
        const int j0 = w << 5;
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
          const int j = j0 + b;
          if (j > i && j < n && iou_gt(bi, B[j], nms_thr)) bits |= 1u << b;
        }
      }
      mat[e] = bits;
    }
    __syncthreads();
    if (tid < 32) {
      // One 32-candidate word at a time, without a shuffle or a branch per candidate: every lane walks the word's
      // greedy chain itself on the word's diagonal block (broadcast loads that do not depend on the chain, so they
      // pipeline; the chain is three ALU operations per candidate), then lane w ORs the rows of the word's keepers
      // into its own word of the removed set.  (A per-candidate loop with a shuffle + a divergent load per step
      // took ~190 cycles per candidate; a variant with one shuffle per keeper was slower still.)
      unsigned int removed = 0;                      // lane w: bits of word w removed by keepers of earlier words
      const int lw = min(tid, nblk - 1);
      for (int wi = 0; wi < nblk; ++wi) {
        const int i0 = wi << 5, left = min(32, n - i0);
        unsigned int alive = ~__shfl_sync(0xffffffffu, removed, wi) & (left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
        unsigned int kept = 0;
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
          const unsigned int d = mat[(i0 + min(b, left - 1)) * nblk + wi];   // bits above b only
          const unsigned int k = (alive >> b) & 1u;
          kept |= k << b;
          alive &= ~(d & (0u - k));
        }
        unsigned int acc = 0;
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
          const unsigned int row = mat[(i0 + min(b, left - 1)) * nblk + lw];
          acc |= row & (0u - ((kept >> b) & 1u));
        }
        if (tid > wi) removed |= acc;
        if (tid == 0) keep[wi] = kept;
      }
    }
    __syncthreads();
  } else {
  // Greedy suppression in sorted order, 32 candidates (one warp) at a time.  Warp 0 walks the blocks: a block
  // whose candidates were all suppressed by earlier keepers costs one ballot; in a block with survivors only
  // the candidates still alive are visited (each may suppress later lanes).  The block's keepers are then
  // published and ALL threads apply them to the rest of the list, so the block-wide barriers are paid once
  // per block that keeps something, not once per 32 candidates.
  int blk = 0;
  while (true) {
    if (tid < 32) {
      unsigned am = 0;
      int bcur = blk;
      for (; bcur < nblk; ++bcur) {
        const int j = (bcur << 5) + tid;
        const bool valid = j < n;
        bool alive = valid && !supp[j];
        unsigned rem = __ballot_sync(0xffffffffu, alive);
        if (rem == 0) {
          if (tid == 0) keep[bcur] = 0;
          continue;
        }
        const int4 bj = valid ? B[j] : make_int4(0, 0, 0, 0);
        unsigned cand = rem;
        while (cand) {
          const int t = __ffs(cand) - 1;               // lowest candidate still alive: it is kept
          int4 bi;
          bi.x = __shfl_sync(0xffffffffu, bj.x, t);
          bi.y = __shfl_sync(0xffffffffu, bj.y, t);
          bi.z = __shfl_sync(0xffffffffu, bj.z, t);
          bi.w = __shfl_sync(0xffffffffu, bj.w, t);
          if (tid > t && alive && iou_gt(bi, bj, nms_thr)) alive = false;
          rem = __ballot_sync(0xffffffffu, alive);
          cand = rem & ~((2u << t) - 1u);              // alive lanes above t
        }
        am = rem;
        if (tid == 0) keep[bcur] = am;
        break;
      }
      if (tid == 0) { blockmask = am; cur_block = bcur; }
    }
    __syncthreads();
    const int bcur = cur_block;
    if (bcur >= nblk) break;
    const unsigned am = blockmask;
    const int base = bcur << 5;
    for (int j = base + 32 + tid; j < n; j += NMS_THREADS) {
      if (supp[j]) continue;
      const int4 bj = B[j];
      unsigned m = am;
      while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        if (iou_gt(B[base + t], bj, nms_thr)) { supp[j] = 1; break; }
      }
    }
    __syncthreads();
    blk = bcur + 1;
  }
  }
  // compaction in sorted order: exclusive prefix over the keep words, then scatter
  if (tid == 0) {
    int run = 0;
    for (int b = 0; b < nblk; ++b) {
      pref[b] = run;
      run += __popc(keep[b]);
    }
    kept_total = run;
  }
  __syncthreads();
  const int kept = kept_total;
  // each kept candidate computes its rank = popcount of earlier keep bits
  for (int i = tid; i < n; i += NMS_THREADS) {
    const int b = i >> 5, l = i & 31;
    const unsigned m = keep[b];
    if (!((m >> l) & 1u)) continue;
    const int rank = pref[b] + __popc(m & ((1u << l) - 1));
    if (rank >= cap) continue;
    const unsigned long long key = K[i];
    const int a = (int)(~(unsigned int)(key & 0xffffffffull));
    int s, local, n_s, ws, stride;
    anchor_locate(a, s, local, n_s, ws, stride);
    const int cell = local >> 1;
    const float cx = (float)((cell % ws) * stride);
    const float cy = (float)((cell / ws) * stride);
    const float st = (float)stride;
    const float* kp = args.h.kps[s] + ((size_t)img * n_s + local) * 10;
    fr_face f;
    const int4 bx = B[i];
    f.x = bx.x; f.y = bx.y; f.w = bx.z; f.h = bx.w;
    f.score = float_unsortable((unsigned int)(key >> 32));
#pragma unroll
    for (int p = 0; p < 5; ++p) {
      f.lm[2 * p] = __fdiv_rn(__fadd_rn(cx, __fmul_rn(__ldg(kp + 2 * p), st)), scale);
      f.lm[2 * p + 1] = __fdiv_rn(__fadd_rn(cy, __fmul_rn(__ldg(kp + 2 * p + 1), st)), scale);
    }
    out[(size_t)img * cap + rank] = f;
  }
  if (tid == 0) n_out[img] = min(kept, cap);
}

constexpr size_t NMS_SMEM = SMEM_CAND * 8 + SMEM_CAND * 16 + ((NA + 15) / 16) * 16 + (NA / 32 + 2) * 8;

}  // namespace

int k_scrfd_decode_nms(fr_ctx* ctx, NmsScratch& s, const HeadPtrs& heads, int n_img,
                       const ImgDesc* d_desc, const float* d_scales, float score_thr,
                       float nms_thr, fr_face* d_out, int cap_per_img, int* d_n_out) {
  if (!s.keys.reserve((size_t)n_img * KEY_STRIDE * 8) || !s.counts.reserve((size_t)n_img * 4) ||
      !s.boxes.reserve((size_t)n_img * NA * 16))
    return fr_fail(ctx, FR_ERR_CUDA, "nms scratch allocation failed");
  FR_CUDA_OK(ctx, fr_opt_in_smem(ctx, nms_kernel, (int)NMS_SMEM));
  FR_CUDA_OK(ctx, cudaMemsetAsync(s.counts.p, 0, (size_t)n_img * 4, ctx->stream));
  DecodeArgs args;
  args.h = heads;
  args.desc = d_desc;
  args.scales = d_scales;
  dim3 grid(ceil_div(NA, 256), n_img);
  decode_threshold_kernel<<<grid, 256, 0, ctx->stream>>>(
      args, score_thr, s.keys.as<unsigned long long>(), s.counts.as<int>());
  nms_kernel<<<n_img, NMS_THREADS, NMS_SMEM, ctx->stream>>>(
      args, nms_thr, s.keys.as<unsigned long long>(), s.boxes.as<int4>(), s.counts.as<int>(),
      d_out, cap_per_img, d_n_out);
  ctx->launches += 2;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}
