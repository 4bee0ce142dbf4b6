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

enum { MODE_IM2COL = 1, MODE_PW = 2 };
constexpr int TP = 128;   // pixels per block tile (flattened over the batch)
constexpr int KC = 16;    // K chunk staged in shared memory

struct ConvArgs {
  const void* in;        // fp32 NCHW (bf16 for the stem)
  float* out;            // fp32 NCHW
  const float* wt;       // [Kpad][Cpad] transposed weights, zero padded (im2col: row = tap*cin + ci)
  const float* b;        // [Cpad]
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

// ---------------------------------------------------------------------------------------
// Tiled SIMT GEMM:  out[co, px] = act( sum_k Wt[k, co] * X[k, px] + b[co] )
//   MODE_PW     : X[ci, px] = in[ci, px]                           (1x1 conv)
//   MODE_IM2COL : X[tap*cin + ci, px] = in[ci, tap of px]          (dense 3x3, pad 1, cin % 16 == 0)
// Block = 128 threads; tile = 128 pixels x (8*CT) channels; thread tile 8 px x CT channels.
// K is consumed in chunks of 16 staged in shared memory; the next chunk's global loads are
// issued into registers before the current chunk is computed (register double buffering), so
// HBM/L2 latency overlaps the FMAs.  All global accesses are coalesced along pixels.
// grid (ceil(total_px/128), cpad/(8*CT)).
template <int CT, int MODE>
__global__ void __launch_bounds__(128)
tile_conv_kernel(ConvArgs a) {
  constexpr int TC = 8 * CT;
  constexpr int WL = (KC * TC + 127) / 128;      // weight loads per thread per chunk
  __shared__ __align__(16) float Xs[KC][TP];
  __shared__ __align__(16) float Ws[KC][TC];
  const int t = threadIdx.x;
  const int hw = a.hout * a.wout;
  const int c0 = blockIdx.y * TC;
  const float* in = reinterpret_cast<const float*>(a.in);
  // this thread's staging pixel
  const int g = blockIdx.x * TP + t;
  const bool gvalid = g < a.total_px;
  const int gn = gvalid ? g / hw : 0;
  const int gp = gvalid ? g - gn * hw : 0;
  const int oy = gp / a.wout, ox = gp - oy * a.wout;
  const size_t plane = (size_t)a.hin * a.win;
  const size_t in_base = (size_t)gn * a.cin * plane;
  const float gmask = gvalid ? 1.f : 0.f;

  float xv[KC], wv[WL];
  auto prefetch = [&](int k0) {
    if (MODE == MODE_PW) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int ci = min(k0 + kk, a.cin - 1);
        xv[kk] = __ldg(in + in_base + (size_t)ci * plane + gp);
      }
    } else {
      // the whole chunk shares one tap (cin % KC == 0)
      const int q = k0 / a.cin, cbase = k0 - q * a.cin;
      const int iy = oy * a.stride - 1 + q / 3, ix = ox * a.stride - 1 + q % 3;
      const bool ok = gvalid && iy >= 0 && iy < a.hin && ix >= 0 && ix < a.win;
      const size_t off = in_base + (size_t)cbase * plane + (ok ? (size_t)iy * a.win + ix : 0);
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) xv[kk] = ok ? __ldg(in + off + (size_t)kk * plane) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < WL; ++i) {
      const int idx = t + i * 128;
      const int kk = idx / TC, c = idx - kk * TC;
      wv[i] = idx < KC * TC ? __ldg(a.wt + (size_t)(k0 + kk) * a.cpad + c0 + c) : 0.f;
    }
  };

  const int cg = t & 7, pg = t >> 3;
  float acc[8][CT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;

  prefetch(0);
  for (int k0 = 0; k0 < a.kpad; k0 += KC) {
#pragma unroll
    for (int kk = 0; kk < KC; ++kk)
      Xs[kk][t] = (MODE == MODE_PW) ? ((k0 + kk < a.cin) ? xv[kk] * gmask : 0.f) : xv[kk];
#pragma unroll
    for (int i = 0; i < WL; ++i) {
      const int idx = t + i * 128;
      if (idx < KC * TC) (&Ws[0][0])[idx] = wv[i];
    }
    __syncthreads();
    if (k0 + KC < a.kpad) prefetch(k0 + KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][pg * 8]);
      const float4 xb = *reinterpret_cast<const float4*>(&Xs[kk][pg * 8 + 4]);
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float w[CT];
      if (CT % 4 == 0) {
#pragma unroll
        for (int j = 0; j < CT; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][cg * CT + j]);
          w[j] = w4.x; w[j + 1] = w4.y; w[j + 2] = w4.z; w[j + 3] = w4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CT; ++j) w[j] = Ws[kk][cg * CT + j];
      }
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
      // channel c of 30: [0,2) score (sigmoid), [2,10) bbox, [10,30) kps; anchor = p*2 + a
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

// ---------------------------------------------------------------------------------------
// Depthwise 3x3 (pad 1, stride 1|2) + bias + ReLU, fp32 NCHW.  One thread = 4 horizontally
// adjacent outputs of one channel (float4 store); memory bound.
// grid (ceil(hout*wout/4/256), c, n).
template <int STRIDE>
__global__ void __launch_bounds__(256)
dw3x3_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
             const float* __restrict__ b, int c_total, int hin, int win, int hout, int wout) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;   // quad index within the plane
  const int qw = wout >> 2;
  if (q >= hout * qw) return;
  const int c = blockIdx.y, n = blockIdx.z;
  const int oy = q / qw, ox0 = (q - oy * qw) * 4;
  const float* ip = in + ((size_t)n * c_total + c) * hin * win;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + c * 9 + i);
  const float bias = __ldg(b + c);
  constexpr int NX = 3 * STRIDE + 3;         // input columns touched: 6 (s1) or 9 (s2)
  const int ix0 = ox0 * STRIDE - 1;
  float acc[4] = {bias, bias, bias, bias};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = oy * STRIDE - 1 + r;
    const bool rok = iy >= 0 && iy < hin;
    float v[NX];
#pragma unroll
    for (int x = 0; x < NX; ++x) {
      const int ix = ix0 + x;
      v[x] = (rok && ix >= 0 && ix < win) ? __ldg(ip + (size_t)iy * win + ix) : 0.f;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int s = 0; s < 3; ++s) acc[o] = fmaf(v[o * STRIDE + s], wk[r * 3 + s], acc[o]);
  }
  float4 res = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
  *reinterpret_cast<float4*>(out + ((size_t)n * c_total + c) * hout * wout + (size_t)oy * wout + ox0) = res;
}

// ---------------------------------------------------------------------------------------
// Stem: dense 3x3 stride-2 conv 3 -> 16 + bias + ReLU on the bf16 planar input produced by K1.
// One thread = one output pixel x 16 channels; weights [27][16] broadcast from shared memory.
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
  const int p = (int)(gid - (size_t)n * HO * HO);
  const int oy = p / HO, ox = p - oy * HO;
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
  float* op = out + (size_t)n * 16 * HO * HO + p;
#pragma unroll
  for (int i = 0; i < 16; ++i) op[(size_t)i * HO * HO] = fmaxf(acc[i], 0.f);
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
  float *a_stem = nullptr, *a_b0 = nullptr, *dw_tmp = nullptr;
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
          int kdim, int cin, bool tap_major = false) {
  pc.cin = cin;
  pc.cout = cout;
  pc.kdim = kdim;
  pc.ct = pick_ct(cout);
  const int tc = 8 * pc.ct;
  pc.cpad = (cout + tc - 1) / tc * tc;
  pc.kpad = (kdim + KC - 1) / KC * KC;
  std::vector<float> wt((size_t)pc.kpad * pc.cpad, 0.f), bp(pc.cpad, 0.f);
  for (int c = 0; c < cout; ++c) {
    for (int k = 0; k < kdim; ++k) {
      // source order is (ci, tap); tap-major kernels want row = tap*cin + ci
      const int row = tap_major ? (k % 9) * cin + k / 9 : k;
      wt[(size_t)row * pc.cpad + c] = w[(size_t)c * kdim + k];
    }
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
  m->dw_tmp = A((size_t)16 * 320 * 320);   // largest depthwise output (b0)
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

template <int CT, int MODE>
int launch_tile(fr_ctx* ctx, const ConvArgs& a) {
  dim3 grid(ceil_div(a.total_px, TP), a.cpad / (8 * CT));
  tile_conv_kernel<CT, MODE><<<grid, 128, 0, ctx->stream>>>(a);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

template <int MODE>
int launch_by_ct(fr_ctx* ctx, const ConvArgs& a, int ct) {
  switch (ct) {
    case 2: return launch_tile<2, MODE>(ctx, a);
    case 4: return launch_tile<4, MODE>(ctx, a);
    case 5: return launch_tile<5, MODE>(ctx, a);
    case 8: return launch_tile<8, MODE>(ctx, a);
    case 9: return launch_tile<9, MODE>(ctx, a);
    case 10: return launch_tile<10, MODE>(ctx, a);
    default: return fr_fail(ctx, FR_ERR_UNSUPPORTED, "unsupported channel tile");
  }
}

int launch_dw(fr_ctx* ctx, const float* in, float* out, const float* w, const float* b, int c, int hin,
              int stride, int n) {
  const int hout = hin / stride;
  dim3 grid(ceil_div(hout * (hout / 4), 256), c, n);
  if (stride == 1) dw3x3_kernel<1><<<grid, 256, 0, ctx->stream>>>(in, out, w, b, c, hin, hin, hout, hout);
  else dw3x3_kernel<2><<<grid, 256, 0, ctx->stream>>>(in, out, w, b, c, hin, hin, hout, hout);
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
    ok = ok && pack(m.get(), m->conv[name], tw.data, w->at(name + ".b").data, cout, cin * k * k, cin,
                    k == 3 && cin % KC == 0);
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
  {
    const fr_tensor& tw = w->at("stem.w");   // [16][3][3][3] -> [27][16]
    std::vector<float> sw(27 * 16);
    for (int co = 0; co < 16; ++co)
      for (int k = 0; k < 27; ++k) sw[k * 16 + co] = tw.data[co * 27 + k];
    PackedConv& pc = m->conv["stem"];
    pc.wt = upload(m.get(), sw);
    pc.b = upload(m.get(), w->at("stem.b").data);
    ok = ok && pc.wt && pc.b;
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
    ok = ok && pack(m.get(), m->conv[h + ".out"], fw, fb, 30, 64 * 9, 64, true);
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
    a.wt = pc.wt; a.b = pc.b;
    a.cin = pc.cin; a.cout = pc.cout; a.cpad = pc.cpad; a.kdim = pc.kdim; a.kpad = pc.kpad;
    a.hin = a.win = hin; a.hout = a.wout = hin / stride; a.stride = stride;
    a.total_px = n * a.hout * a.wout;
    return a;
  };
  // stem: 3x3 s2, 3 -> 16, ReLU (bf16 planar input from K1)
  {
    const PackedConv& pc = m->conv.at("stem");
    const size_t total = (size_t)n * (DET / 2) * (DET / 2);
    stem_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_in_chw, m->a_stem, pc.wt, pc.b, n);
    ctx->launches++;
    FR_CUDA_OK(ctx, cudaGetLastError());
  }
  // depthwise-separable block: dw3x3(stride)+ReLU into the scratch plane set, then 1x1+ReLU
  auto dwsep = [&](const std::string& name, const float* in, float* out, int hin, int stride) -> int {
    const PackedConv& pc = m->conv.at(name);
    FR_CHECK(launch_dw(ctx, in, m->dw_tmp, pc.wd, pc.bd, pc.cin, hin, stride, n));
    ConvArgs a = args(pc, m->dw_tmp, out, hin / stride, 1);
    a.relu = 1;
    return launch_by_ct<MODE_PW>(ctx, a, pc.ct);
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
    FR_CHECK((launch_by_ct<MODE_PW>(ctx, a, pc.ct)));
  }
  auto conv3 = [&](const std::string& name, const float* in, float* out, int hin, int stride,
                   int accumulate) -> int {
    const PackedConv& pc = m->conv.at(name);
    ConvArgs a = args(pc, in, out, hin, stride);
    a.accumulate = accumulate;
    return launch_by_ct<MODE_IM2COL>(ctx, a, pc.ct);
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
    FR_CHECK((launch_by_ct<MODE_IM2COL>(ctx, a, pc.ct)));
    heads->score[i] = m->score[i];
    heads->bbox[i] = m->bbox[i];
    heads->kps[i] = m->kps[i];
  }
  return FR_OK;
}
