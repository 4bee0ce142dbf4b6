// K2: SCRFD det_500m forward -- the replacement for `session_->Run` in FaceDetector::detect
// (reference src/face_detector.cpp:170-183).  Architecture per InsightFace scrfd_500m_bnkps
// (depthwise-separable backbone 16/40/72/152/288, PAFPN 16 ch, per-stride heads 64 ch,
// 2 anchors), BN folded into conv+bias, sigmoid on the score heads.
//
// The network is bandwidth / latency bound (depthwise 3x3 + thin 1x1; 1.47 GFLOP per frame,
// channel counts hostile to MMA tiles) and must hold boxes to 1e-3 px, so it runs in fp32 on
// the CUDA cores: NCHW planar activations, one thread per output pixel, weights broadcast
// from shared memory as float4, depthwise 3x3 + ReLU fused into the following pointwise conv
// (the depthwise value is recomputed per output-channel chunk instead of round-tripping HBM).
#include <cmath>
#include <cstring>

#include "common.h"

namespace {

constexpr int DET = FR_DET_SIZE;

enum { MODE_DW = 0, MODE_IM2COL = 1, MODE_PW = 2 };
constexpr int TP = 128;   // pixels per block tile (flattened over the batch)
constexpr int KC = 16;    // K chunk staged in shared memory

struct ConvArgs {
  const void* in;        // fp32 NCHW (or bf16 for the stem)
  float* out;            // fp32 NCHW
  const float* wt;       // [Kpad][Cpad] transposed weights (K = cin, or cin*9 for im2col), zero padded
  const float* b;        // [Cpad]
  const float* wd;       // depthwise weights [cin][9] (MODE_DW)
  const float* bd;       // depthwise bias [cin]
  const float* add_up;   // optional [n][cout][hout/2][wout/2], nearest-upsampled and added
  int cin, cout, cpad, kdim, kpad;
  int hin, win, hout, wout, stride;
  int total_px;          // n * hout * wout
  int relu;
  int accumulate;        // out += result (PAFPN bottom-up path)
  int head;              // scatter to score/bbox/kps (anchor-major) with sigmoid on the scores
  float* score;
  float* bbox;
  float* kps;
};

// Tiled SIMT GEMM:  out[co, px] = act( sum_k Wt[k, co] * X[k, px] + b[co] )
//   MODE_DW     : X[ci, px] = relu(depthwise3x3(in)[ci, px] + bd[ci])      (fused dw-separable)
//   MODE_IM2COL : X[ci*9+t, px] = in[ci, tap t of px]                      (dense 3x3, pad 1)
//   MODE_PW     : X[ci, px] = in[ci, px]                                   (1x1)
// Block = 128 threads; tile = 128 pixels x (8*CT) output channels; thread tile 8 px x CT co.
// X and W chunks (KC deep) are staged in shared memory; all global reads are coalesced.
// grid (ceil(total_px/128), cpad/(8*CT)).
template <int CT, int MODE, bool IN_BF16>
__global__ void __launch_bounds__(128)
tile_conv_kernel(ConvArgs a) {
  constexpr int TC = 8 * CT;
  __shared__ __align__(16) float Xs[KC][TP];
  __shared__ __align__(16) float Ws[KC][TC];
  extern __shared__ __align__(16) float dwsm[];  // MODE_DW: wd[cin][9] | bd[cin]
  const int t = threadIdx.x;
  const int hw = a.hout * a.wout;
  const int c0 = blockIdx.y * TC;
  if (MODE == MODE_DW) {
    for (int i = t; i < a.cin * 9; i += 128) dwsm[i] = a.wd[i];
    for (int i = t; i < a.cin; i += 128) dwsm[a.cin * 9 + i] = a.bd[i];
    __syncthreads();
  }
  // this thread's staging pixel
  const int g = blockIdx.x * TP + t;
  const bool gvalid = g < a.total_px;
  const int gn = gvalid ? g / hw : 0;
  const int gp = gvalid ? g - gn * hw : 0;
  const int oy = gp / a.wout, ox = gp - oy * a.wout;
  const int iy0 = oy * a.stride - 1, ix0 = ox * a.stride - 1;
  const size_t plane = (size_t)a.hin * a.win;
  const size_t in_base = (size_t)gn * a.cin * plane;
  auto ld = [&](size_t idx) -> float {
    if (IN_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.in)[idx]);
    return __ldg(reinterpret_cast<const float*>(a.in) + idx);
  };
  // 3x3 tap validity / offsets (shared by every channel)
  int toff[9];
  bool tok[9];
  if (MODE != MODE_PW) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = iy0 + k / 3, ix = ix0 + k % 3;
      tok[k] = gvalid && iy >= 0 && iy < a.hin && ix >= 0 && ix < a.win;
      toff[k] = iy * a.win + ix;
    }
  }
  const int cg = t & 7, pg = t >> 3;
  float acc[8][CT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.kpad; k0 += KC) {
    // ---- stage X[k0..k0+KC) for this thread's pixel
    if (MODE == MODE_DW) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int ci = k0 + kk;
        float v = 0.f;
        if (ci < a.cin) {
          const float* wdp = dwsm + ci * 9;
          const size_t cb = in_base + (size_t)ci * plane;
          v = dwsm[a.cin * 9 + ci];
#pragma unroll
          for (int q = 0; q < 9; ++q)
            if (tok[q]) v = fmaf(ld(cb + toff[q]), wdp[q], v);
          v = fmaxf(v, 0.f);
        }
        Xs[kk][t] = v;
      }
    } else if (MODE == MODE_IM2COL) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int k = k0 + kk;
        const int ci = k / 9, q = k - ci * 9;
        float v = 0.f;
        if (k < a.kdim) {
          // tok/toff indexed dynamically -> select through a small unrolled chain
          bool ok = false;
          int off = 0;
#pragma unroll
          for (int r = 0; r < 9; ++r)
            if (r == q) { ok = tok[r]; off = toff[r]; }
          if (ok) v = ld(in_base + (size_t)ci * plane + off);
        }
        Xs[kk][t] = v;
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int ci = k0 + kk;
        Xs[kk][t] = (gvalid && ci < a.cin) ? ld(in_base + (size_t)ci * plane + gp) : 0.f;
      }
    }
    // ---- stage W[k0..k0+KC)[c0..c0+TC) (coalesced; padded so no guards are needed)
    for (int i = t; i < KC * TC; i += 128) {
      const int kk = i / TC, c = i - kk * TC;
      Ws[kk][c] = __ldg(a.wt + (size_t)(k0 + kk) * a.cpad + c0 + c);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][pg * 8]);
      const float4 xb = *reinterpret_cast<const float4*>(&Xs[kk][pg * 8 + 4]);
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float w[CT];
#pragma unroll
      for (int j = 0; j < CT; ++j) w[j] = Ws[kk][cg * CT + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue: this thread owns pixels g0..g0+7 (same image: hw % 8 == 0) x CT channels
  const int g0 = blockIdx.x * TP + pg * 8;
  if (g0 >= a.total_px) return;
  const int n = g0 / hw;
  const int p0 = g0 - n * hw;
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const int c = c0 + cg * CT + j;
    if (c >= a.cout) continue;
    const float bias = __ldg(a.b + c);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = acc[i][j] + bias;
      if (a.relu) v[i] = fmaxf(v[i], 0.f);
    }
    if (a.head) {
      // channel c of 30: [0,2) score (sigmoid), [2,10) bbox, [10,30) kps; anchor = p*2 + (c % 2 for score)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const size_t p = (size_t)p0 + i;
        if (c < 2) a.score[(size_t)n * hw * 2 + p * 2 + c] = 1.0f / (1.0f + expf(-v[i]));
        else if (c < 10) a.bbox[(size_t)n * hw * 8 + p * 8 + (c - 2)] = v[i];
        else a.kps[(size_t)n * hw * 20 + p * 20 + (c - 10)] = v[i];
      }
      continue;
    }
    if (a.add_up) {
      const int uw = a.wout >> 1, uh = a.hout >> 1;
      const float* up = a.add_up + ((size_t)n * a.cout + c) * uh * uw;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = p0 + i;
        const int y = p / a.wout, x = p - y * a.wout;
        v[i] += __ldg(up + (size_t)(y >> 1) * uw + (x >> 1));
      }
    }
    float4* o = reinterpret_cast<float4*>(a.out + ((size_t)n * a.cout + c) * hw + p0);
    if (a.accumulate) {
      const float4 o0 = o[0], o1 = o[1];
      v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
      v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
    }
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

}  // namespace

struct PackedConv {       // device-side packed parameters of one (fused) conv
  float* wt = nullptr;     // [kpad][cpad]
  float* b = nullptr;      // [cpad]
  float* wd = nullptr;     // [cin][9]   (dw-separable only)
  float* bd = nullptr;     // [cin]
  int cin = 0, cout = 0, cpad = 0, kdim = 0, kpad = 0, ct = 2;
};

struct DetModel {
  std::map<std::string, PackedConv> conv;
  std::vector<void*> allocs;
  int cap = 0;
  std::vector<void*> act_allocs;
  // activations (fp32 NCHW)
  float *a_stem = nullptr, *a_b0 = nullptr;
  std::vector<float*> a_stage;       // per dwsep block output
  float* lat[3] = {nullptr, nullptr, nullptr};
  float* inter[3] = {nullptr, nullptr, nullptr};
  float* pout[3] = {nullptr, nullptr, nullptr};
  float* tw0[3] = {nullptr, nullptr, nullptr};
  float* tw1[3] = {nullptr, nullptr, nullptr};
  float* score[3] = {nullptr, nullptr, nullptr};
  float* bbox[3] = {nullptr, nullptr, nullptr};
  float* kps[3] = {nullptr, nullptr, nullptr};
};

namespace {

const int kStages[4][2] = {{2, 40}, {3, 72}, {2, 152}, {6, 288}};

// output-channel thread tile: tile width is 8*CT channels
int pick_ct(int cout) {
  switch (cout) {
    case 16: return 2;
    case 30: return 4;
    case 40: return 5;
    case 64: return 8;
    case 72: return 9;
    case 152: return 10;
    case 288: return 9;
    default: return 4;
  }
}

float* upload(DetModel* m, const std::vector<float>& h) {
  float* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  m->allocs.push_back(d);
  return d;
}

// w: [cout][kdim] row-major (OIHW flattened) -> transposed, zero-padded [kpad][cpad]
bool pack(DetModel* m, PackedConv& pc, const std::vector<float>& w, const std::vector<float>& b, int cout,
          int kdim, int cin) {
  pc.cin = cin;
  pc.cout = cout;
  pc.kdim = kdim;
  pc.ct = pick_ct(cout);
  const int tc = 8 * pc.ct;
  pc.cpad = (cout + tc - 1) / tc * tc;
  pc.kpad = (kdim + KC - 1) / KC * KC;
  std::vector<float> wt((size_t)pc.kpad * pc.cpad, 0.f), bp(pc.cpad, 0.f);
  for (int c = 0; c < cout; ++c) {
    for (int k = 0; k < kdim; ++k) wt[(size_t)k * pc.cpad + c] = w[(size_t)c * kdim + k];
    bp[c] = b[c];
  }
  pc.wt = upload(m, wt);
  pc.b = upload(m, bp);
  return pc.wt && pc.b;
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

template <int CT, int MODE, bool BF>
int launch_tile(fr_ctx* ctx, const ConvArgs& a) {
  const size_t dsm = MODE == MODE_DW ? (size_t)a.cin * 10 * sizeof(float) : 0;
  dim3 grid(ceil_div(a.total_px, TP), a.cpad / (8 * CT));
  tile_conv_kernel<CT, MODE, BF><<<grid, 128, dsm, ctx->stream>>>(a);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

template <int MODE, bool BF>
int launch_by_ct(fr_ctx* ctx, const ConvArgs& a, int ct) {
  switch (ct) {
    case 2: return launch_tile<2, MODE, BF>(ctx, a);
    case 4: return launch_tile<4, MODE, BF>(ctx, a);
    case 5: return launch_tile<5, MODE, BF>(ctx, a);
    case 8: return launch_tile<8, MODE, BF>(ctx, a);
    case 9: return launch_tile<9, MODE, BF>(ctx, a);
    case 10: return launch_tile<10, MODE, BF>(ctx, a);
    default: return fr_fail(ctx, FR_ERR_UNSUPPORTED, "unsupported channel tile");
  }
}

}  // namespace

int det_model_create(fr_ctx* ctx, const fr_weights* w) {
  if (!w || w->model != FR_MODEL_DET) return fr_fail(ctx, FR_ERR_MODEL, "det weights missing");
  std::unique_ptr<DetModel> m(new DetModel());
  bool ok = true;
  auto dense = [&](const std::string& name) {   // conv (3x3 or 1x1) as a [cout][cin*k*k] GEMM
    const fr_tensor& tw = w->at(name + ".w");
    const int cout = (int)tw.dims[0], cin = (int)tw.dims[1], k = (int)tw.dims[2];
    ok = ok && pack(m.get(), m->conv[name], tw.data, w->at(name + ".b").data, cout, cin * k * k, cin);
  };
  auto dwsep = [&](const std::string& name) {   // dw 3x3 (+ReLU) fused in front of the 1x1
    const fr_tensor& pw = w->at(name + ".pw.w");
    const int cout = (int)pw.dims[0], cin = (int)pw.dims[1];
    PackedConv& pc = m->conv[name];
    ok = ok && pack(m.get(), pc, pw.data, w->at(name + ".pw.b").data, cout, cin, cin);
    pc.wd = upload(m.get(), w->at(name + ".dw.w").data);
    pc.bd = upload(m.get(), w->at(name + ".dw.b").data);
    ok = ok && pc.wd && pc.bd;
  };
  dense("stem");
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
    ok = ok && pack(m.get(), m->conv[h + ".out"], fw, fb, 30, 64 * 9, 64);
  }
  if (!ok) {
    for (void* p : m->allocs) cudaFree(p);
    return fr_fail(ctx, FR_ERR_CUDA, "det weight upload failed");
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
  auto args = [&](const PackedConv& pc, const void* in, float* out, int hin, int stride) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.out = out;
    a.wt = pc.wt; a.b = pc.b; a.wd = pc.wd; a.bd = pc.bd;
    a.cin = pc.cin; a.cout = pc.cout; a.cpad = pc.cpad; a.kdim = pc.kdim; a.kpad = pc.kpad;
    a.hin = a.win = hin; a.hout = a.wout = hin / stride; a.stride = stride;
    a.total_px = n * a.hout * a.wout;
    return a;
  };
  // stem: 3x3 s2, 3 -> 16, ReLU (bf16 planar input from K1)
  {
    const PackedConv& pc = m->conv.at("stem");
    ConvArgs a = args(pc, d_in_chw, m->a_stem, DET, 2);
    a.relu = 1;
    FR_CHECK((launch_by_ct<MODE_IM2COL, true>(ctx, a, pc.ct)));
  }
  auto dwsep = [&](const std::string& name, const float* in, float* out, int hin, int stride) -> int {
    const PackedConv& pc = m->conv.at(name);
    ConvArgs a = args(pc, in, out, hin, stride);
    a.relu = 1;
    return launch_by_ct<MODE_DW, false>(ctx, a, pc.ct);
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
  // laterals (1x1, no activation) with the top-down nearest-2x add fused in
  for (int i = 2; i >= 0; --i) {
    const PackedConv& pc = m->conv.at("lat" + std::to_string(i));
    ConvArgs a = args(pc, feats[i], m->lat[i], fh[i], 1);
    a.add_up = i < 2 ? m->lat[i + 1] : nullptr;
    FR_CHECK((launch_by_ct<MODE_PW, false>(ctx, a, pc.ct)));
  }
  auto conv3 = [&](const std::string& name, const float* in, float* out, int hin, int stride,
                   int accumulate) -> int {
    const PackedConv& pc = m->conv.at(name);
    ConvArgs a = args(pc, in, out, hin, stride);
    a.accumulate = accumulate;
    return launch_by_ct<MODE_IM2COL, false>(ctx, a, pc.ct);
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
    const PackedConv& pc = m->conv.at(h + ".out");
    ConvArgs a = args(pc, m->tw1[i], nullptr, fh[i], 1);
    a.head = 1; a.score = m->score[i]; a.bbox = m->bbox[i]; a.kps = m->kps[i];
    FR_CHECK((launch_by_ct<MODE_IM2COL, false>(ctx, a, pc.ct)));
    heads->score[i] = m->score[i];
    heads->bbox[i] = m->bbox[i];
    heads->kps[i] = m->kps[i];
  }
  return FR_OK;
}
