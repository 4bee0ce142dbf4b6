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

struct ConvArgs {
  const void* in;        // fp32 NCHW (or bf16 for the stem)
  float* out;            // fp32 NCHW
  const float* w;        // packed (see kernels)
  const float* b;
  const float* wd;       // depthwise weights [cin][9] (dwpw only)
  const float* bd;       // depthwise bias [cin]
  const float* add_up;   // optional [n][cout][Hout/2][Wout/2] nearest-upsampled and added
  int cin, cout;
  int hin, win, hout, wout, stride;
  int relu;
  int accumulate;        // out += result (PAFPN bottom-up path)
  // head scatter (mode 2): score/bbox/kps in anchor-major layout, sigmoid on scores
  float* score;
  float* bbox;
  float* kps;
  int head;
};

// Dense 3x3 conv, pad 1.  grid (ceil(hout*wout/128), ceil(cout/CO_T), n).
// smem: w[cin*9][CO_T] | b[CO_T]
template <int CO_T, bool IN_BF16>
__global__ void __launch_bounds__(128)
conv3x3_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;
  float* sb = sm + (size_t)a.cin * 9 * CO_T;
  const int co0 = blockIdx.y * CO_T;
  for (int i = threadIdx.x; i < a.cin * 9 * CO_T; i += blockDim.x) {
    const int co = i % CO_T, k = i / CO_T;  // k = ci*9 + t
    sw[i] = (co0 + co < a.cout) ? a.w[(size_t)(co0 + co) * a.cin * 9 + k] : 0.f;
  }
  if (threadIdx.x < CO_T) sb[threadIdx.x] = (co0 + threadIdx.x < a.cout) ? a.b[co0 + threadIdx.x] : 0.f;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.hout * a.wout) return;
  const int n = blockIdx.z;
  const int oy = p / a.wout, ox = p % a.wout;
  const int iy0 = oy * a.stride - 1, ix0 = ox * a.stride - 1;
  float acc[CO_T];
#pragma unroll
  for (int i = 0; i < CO_T; ++i) acc[i] = sb[i];
  const size_t plane = (size_t)a.hin * a.win;
  for (int ci = 0; ci < a.cin; ++ci) {
    const size_t base = ((size_t)n * a.cin + ci) * plane;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int iy = iy0 + t / 3, ix = ix0 + t % 3;
      float v = 0.f;
      if (iy >= 0 && iy < a.hin && ix >= 0 && ix < a.win) {
        const size_t idx = base + (size_t)iy * a.win + ix;
        if (IN_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.in)[idx]);
        else v = __ldg(reinterpret_cast<const float*>(a.in) + idx);
      }
      const float4* wp = reinterpret_cast<const float4*>(sw + (size_t)(ci * 9 + t) * CO_T);
#pragma unroll
      for (int i = 0; i < CO_T / 4; ++i) {
        const float4 w4 = wp[i];
        acc[4 * i] = fmaf(v, w4.x, acc[4 * i]);
        acc[4 * i + 1] = fmaf(v, w4.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(v, w4.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(v, w4.w, acc[4 * i + 3]);
      }
    }
  }
  if (a.head) {
    // channel c of 30: [0,2) score (sigmoid), [2,10) bbox, [10,30) kps; anchor = p*2 + a
    const size_t hw = (size_t)a.hout * a.wout;
#pragma unroll
    for (int i = 0; i < CO_T; ++i) {
      const int c = co0 + i;
      if (c >= a.cout) break;
      const float v = acc[i];
      if (c < 2) a.score[(size_t)n * hw * 2 + (size_t)p * 2 + c] = 1.0f / (1.0f + expf(-v));
      else if (c < 10) a.bbox[(size_t)n * hw * 8 + (size_t)p * 8 + (c - 2)] = v;
      else a.kps[(size_t)n * hw * 20 + (size_t)p * 20 + (c - 10)] = v;
    }
    return;
  }
  const size_t oplane = (size_t)a.hout * a.wout;
#pragma unroll
  for (int i = 0; i < CO_T; ++i) {
    const int c = co0 + i;
    if (c >= a.cout) break;
    float v = acc[i];
    if (a.relu) v = fmaxf(v, 0.f);
    float* o = a.out + ((size_t)n * a.cout + c) * oplane + p;
    if (a.accumulate) v += *o;
    *o = v;
  }
}

// [optional depthwise 3x3 (stride s) + ReLU] + pointwise 1x1 + bias [+ ReLU] [+ upsampled add].
// grid (ceil(hout*wout/128), ceil(cout/CO_T), n).
// smem: wp[cin][CO_T] | bp[CO_T] | wd[cin][9] | bd[cin]
template <int CO_T, bool DW>
__global__ void __launch_bounds__(128)
dwpw_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* swp = sm;
  float* sbp = swp + (size_t)a.cin * CO_T;
  float* swd = sbp + CO_T;
  float* sbd = swd + (size_t)a.cin * 9;
  const int co0 = blockIdx.y * CO_T;
  for (int i = threadIdx.x; i < a.cin * CO_T; i += blockDim.x) {
    const int co = i % CO_T, ci = i / CO_T;
    swp[i] = (co0 + co < a.cout) ? a.w[(size_t)(co0 + co) * a.cin + ci] : 0.f;
  }
  if (threadIdx.x < CO_T) sbp[threadIdx.x] = (co0 + threadIdx.x < a.cout) ? a.b[co0 + threadIdx.x] : 0.f;
  if (DW) {
    for (int i = threadIdx.x; i < a.cin * 9; i += blockDim.x) swd[i] = a.wd[i];
    for (int i = threadIdx.x; i < a.cin; i += blockDim.x) sbd[i] = a.bd[i];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.hout * a.wout) return;
  const int n = blockIdx.z;
  const int oy = p / a.wout, ox = p % a.wout;
  const int iy0 = oy * a.stride - 1, ix0 = ox * a.stride - 1;
  float acc[CO_T];
#pragma unroll
  for (int i = 0; i < CO_T; ++i) acc[i] = sbp[i];
  const size_t plane = (size_t)a.hin * a.win;
  const float* in = reinterpret_cast<const float*>(a.in) + (size_t)n * a.cin * plane;
  for (int ci = 0; ci < a.cin; ++ci) {
    float v;
    if (DW) {
      v = sbd[ci];
      const float* ip = in + (size_t)ci * plane;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int iy = iy0 + t / 3, ix = ix0 + t % 3;
        if (iy >= 0 && iy < a.hin && ix >= 0 && ix < a.win)
          v = fmaf(__ldg(ip + (size_t)iy * a.win + ix), swd[ci * 9 + t], v);
      }
      v = fmaxf(v, 0.f);
    } else {
      v = __ldg(in + (size_t)ci * plane + p);
    }
    const float4* wp = reinterpret_cast<const float4*>(swp + (size_t)ci * CO_T);
#pragma unroll
    for (int i = 0; i < CO_T / 4; ++i) {
      const float4 w4 = wp[i];
      acc[4 * i] = fmaf(v, w4.x, acc[4 * i]);
      acc[4 * i + 1] = fmaf(v, w4.y, acc[4 * i + 1]);
      acc[4 * i + 2] = fmaf(v, w4.z, acc[4 * i + 2]);
      acc[4 * i + 3] = fmaf(v, w4.w, acc[4 * i + 3]);
    }
  }
  const size_t oplane = (size_t)a.hout * a.wout;
#pragma unroll
  for (int i = 0; i < CO_T; ++i) {
    const int c = co0 + i;
    if (c >= a.cout) break;
    float v = acc[i];
    if (a.relu) v = fmaxf(v, 0.f);
    if (a.add_up) {
      const int uw = a.wout >> 1, uh = a.hout >> 1;
      v += __ldg(a.add_up + ((size_t)n * a.cout + c) * uh * uw + (size_t)(oy >> 1) * uw + (ox >> 1));
    }
    a.out[((size_t)n * a.cout + c) * oplane + p] = v;
  }
}

}  // namespace

struct DetModel {
  std::map<std::string, float*> t;   // device copies of every canonical tensor (OIHW as-is)
  std::map<std::string, float*> fused;  // head cls/reg/kps concatenated
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

template <int CO_T, bool BF>
int launch_conv3(fr_ctx* ctx, const ConvArgs& a, int n) {
  const size_t smem = ((size_t)a.cin * 9 * CO_T + CO_T) * sizeof(float);
  dim3 grid(ceil_div(a.hout * a.wout, 128), ceil_div(a.cout, CO_T), n);
  conv3x3_kernel<CO_T, BF><<<grid, 128, smem, ctx->stream>>>(a);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

template <int CO_T, bool DW>
int launch_dwpw(fr_ctx* ctx, const ConvArgs& a, int n) {
  const size_t smem = ((size_t)a.cin * CO_T + CO_T + (size_t)a.cin * 9 + a.cin) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    FR_CUDA_OK(ctx, cudaFuncSetAttribute(dwpw_kernel<CO_T, DW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  dim3 grid(ceil_div(a.hout * a.wout, 128), ceil_div(a.cout, CO_T), n);
  dwpw_kernel<CO_T, DW><<<grid, 128, smem, ctx->stream>>>(a);
  ctx->launches++;
  FR_CUDA_OK(ctx, cudaGetLastError());
  return FR_OK;
}

float* upload(fr_ctx* ctx, DetModel* m, const std::vector<float>& h) {
  float* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  m->allocs.push_back(d);
  return d;
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

}  // namespace

int det_model_create(fr_ctx* ctx, const fr_weights* w) {
  if (!w || w->model != FR_MODEL_DET) return fr_fail(ctx, FR_ERR_MODEL, "det weights missing");
  std::unique_ptr<DetModel> m(new DetModel());
  for (const fr_tensor& t : w->tensors) {
    float* d = upload(ctx, m.get(), t.data);
    if (!d) return fr_fail(ctx, FR_ERR_CUDA, "det weight upload failed");
    m->t[t.name] = d;
  }
  // fuse the three head convs of each stride into one 64 -> 30 conv (cls 2 | reg 8 | kps 20)
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    std::vector<float> fw, fb;
    for (const char* part : {".cls", ".reg", ".kps"}) {
      const fr_tensor& tw = w->at(h + part + ".w");
      const fr_tensor& tb = w->at(h + part + ".b");
      fw.insert(fw.end(), tw.data.begin(), tw.data.end());
      fb.insert(fb.end(), tb.data.begin(), tb.data.end());
    }
    m->fused[h + ".w"] = upload(ctx, m.get(), fw);
    m->fused[h + ".b"] = upload(ctx, m.get(), fb);
    if (!m->fused[h + ".w"] || !m->fused[h + ".b"]) return fr_fail(ctx, FR_ERR_CUDA, "det weight upload failed");
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
  auto W = [&](const std::string& name) { return m->t.at(name); };
  auto base_args = [&]() {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.stride = 1;
    return a;
  };
  // stem: 3x3 s2, 3 -> 16, ReLU (bf16 planar input from K1)
  {
    ConvArgs a = base_args();
    a.in = d_in_chw; a.out = m->a_stem; a.w = W("stem.w"); a.b = W("stem.b");
    a.cin = 3; a.cout = 16; a.hin = a.win = DET; a.hout = a.wout = DET / 2; a.stride = 2; a.relu = 1;
    FR_CHECK((launch_conv3<16, true>(ctx, a, n)));
  }
  auto dwsep = [&](const std::string& name, const float* in, float* out, int cin, int cout, int hin,
                   int stride) -> int {
    ConvArgs a = base_args();
    a.in = in; a.out = out;
    a.w = W(name + ".pw.w"); a.b = W(name + ".pw.b");
    a.wd = W(name + ".dw.w"); a.bd = W(name + ".dw.b");
    a.cin = cin; a.cout = cout; a.hin = a.win = hin; a.hout = a.wout = hin / stride;
    a.stride = stride; a.relu = 1;
    if (cout <= 16) return launch_dwpw<16, true>(ctx, a, n);
    return launch_dwpw<32, true>(ctx, a, n);
  };
  FR_CHECK(dwsep("b0", m->a_stem, m->a_b0, 16, 16, 320, 1));
  const float* cur = m->a_b0;
  int cin = 16, hw = 320, bi = 0;
  const float* feats[3] = {nullptr, nullptr, nullptr};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < kStages[s][0]; ++b, ++bi) {
      const int stride = b == 0 ? 2 : 1;
      FR_CHECK(dwsep("s" + std::to_string(s) + "." + std::to_string(b), cur, m->a_stage[bi], cin,
                     kStages[s][1], hw, stride));
      hw /= stride;
      cin = kStages[s][1];
      cur = m->a_stage[bi];
    }
    if (s >= 1) feats[s - 1] = cur;
  }
  const int fc[3] = {72, 152, 288};
  const int fh[3] = {80, 40, 20};
  // laterals (1x1, no activation) with the top-down nearest-2x add fused in
  for (int i = 2; i >= 0; --i) {
    ConvArgs a = base_args();
    a.in = feats[i]; a.out = m->lat[i];
    a.w = W("lat" + std::to_string(i) + ".w"); a.b = W("lat" + std::to_string(i) + ".b");
    a.cin = fc[i]; a.cout = 16; a.hin = a.win = a.hout = a.wout = fh[i];
    a.add_up = i < 2 ? m->lat[i + 1] : nullptr;
    FR_CHECK((launch_dwpw<16, false>(ctx, a, n)));
  }
  auto conv3 = [&](const std::string& name, const float* in, float* out, int hin, int stride,
                   int accumulate) -> int {
    ConvArgs a = base_args();
    a.in = in; a.out = out; a.w = W(name + ".w"); a.b = W(name + ".b");
    a.cin = 16; a.cout = 16; a.hin = a.win = hin; a.hout = a.wout = hin / stride;
    a.stride = stride; a.accumulate = accumulate;
    return launch_conv3<16, false>(ctx, a, n);
  };
  for (int i = 0; i < 3; ++i) FR_CHECK(conv3("fpn" + std::to_string(i), m->lat[i], m->inter[i], fh[i], 1, 0));
  for (int i = 0; i < 2; ++i)
    FR_CHECK(conv3("down" + std::to_string(i), m->inter[i], m->inter[i + 1], fh[i], 2, 1));
  const float* outs[3] = {m->inter[0], m->pout[1], m->pout[2]};
  for (int i = 1; i < 3; ++i)
    FR_CHECK(conv3("pafpn" + std::to_string(i - 1), m->inter[i], m->pout[i], fh[i], 1, 0));
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    FR_CHECK(dwsep(h + ".t0", outs[i], m->tw0[i], 16, 64, fh[i], 1));
    FR_CHECK(dwsep(h + ".t1", m->tw0[i], m->tw1[i], 64, 64, fh[i], 1));
    ConvArgs a = base_args();
    a.in = m->tw1[i]; a.w = m->fused.at(h + ".w"); a.b = m->fused.at(h + ".b");
    a.cin = 64; a.cout = 30; a.hin = a.win = a.hout = a.wout = fh[i];
    a.head = 1; a.score = m->score[i]; a.bbox = m->bbox[i]; a.kps = m->kps[i];
    FR_CHECK((launch_conv3<16, false>(ctx, a, n)));
    heads->score[i] = m->score[i];
    heads->bbox[i] = m->bbox[i];
    heads->kps[i] = m->kps[i];
  }
  return FR_OK;
}
