// Host-only weight store: the canonical tensor list of the two fixed architectures, seeded
// random initialisation, and the fr_weights_* C ABI.
//
// Replaces what FaceDetector::loadModel / FaceRecognizer::loadModel obtain from the .onnx
// files (reference src/face_detector.cpp:20-90, src/face_recognizer.cpp:21-91).  The
// architectures themselves are not in the reference (they live in the model files);
// they follow InsightFace's scrfd_500m_bnkps and arcface_torch iresnet50 in the exported
// parameterisation (conv+BN folded; pre-conv BN kept as an affine).  The same tensor list is
// restated in oracle/nets.py and compared by tests/test_capi_host.py.
#include <cmath>
#include <cstring>
#include <fstream>

#include "common.h"

static thread_local std::string g_werr;

const fr_tensor& fr_weights::at(const std::string& n) const {
  auto it = index.find(n);
  if (it == index.end()) {
    fprintf(stderr, "fr_weights: missing tensor %s\n", n.c_str());
    abort();
  }
  return tensors[it->second];
}

namespace {

void add(fr_weights& w, const std::string& name, std::vector<int64_t> dims) {
  fr_tensor t;
  t.name = name;
  t.dims = std::move(dims);
  t.data.assign(t.numel(), 0.f);
  w.index[name] = (int)w.tensors.size();
  w.tensors.push_back(std::move(t));
}
void add_conv(fr_weights& w, const std::string& n, int co, int ci, int k) {
  add(w, n + ".w", {co, ci, k, k});
  add(w, n + ".b", {co});
}
void add_dwsep(fr_weights& w, const std::string& n, int ci, int co) {
  add_conv(w, n + ".dw", ci, 1, 3);
  add_conv(w, n + ".pw", co, ci, 1);
}

const int kDetStages[4][2] = {{2, 40}, {3, 72}, {2, 152}, {6, 288}};
const int kRecLayers[4][2] = {{3, 64}, {4, 128}, {14, 256}, {3, 512}};

void build_det(fr_weights& w) {
  add_conv(w, "stem", 16, 3, 3);
  add_dwsep(w, "b0", 16, 16);
  int cin = 16;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kDetStages[s][0]; ++b) {
      add_dwsep(w, "s" + std::to_string(s) + "." + std::to_string(b), cin, kDetStages[s][1]);
      cin = kDetStages[s][1];
    }
  const int feats[3] = {72, 152, 288};
  for (int i = 0; i < 3; ++i) add_conv(w, "lat" + std::to_string(i), 16, feats[i], 1);
  for (int i = 0; i < 3; ++i) add_conv(w, "fpn" + std::to_string(i), 16, 16, 3);
  for (int i = 0; i < 2; ++i) add_conv(w, "down" + std::to_string(i), 16, 16, 3);
  for (int i = 0; i < 2; ++i) add_conv(w, "pafpn" + std::to_string(i), 16, 16, 3);
  for (int i = 0; i < 3; ++i) {
    std::string h = "h" + std::to_string(i);
    add_dwsep(w, h + ".t0", 16, 64);
    add_dwsep(w, h + ".t1", 64, 64);
    add_conv(w, h + ".cls", 2, 64, 3);
    add_conv(w, h + ".reg", 8, 64, 3);
    add_conv(w, h + ".kps", 20, 64, 3);
  }
}

void build_rec(fr_weights& w) {
  add(w, "stem.w", {64, 3, 3, 3});
  add(w, "stem.b", {64});
  add(w, "stem.prelu", {64});
  int cin = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < kRecLayers[l][0]; ++b) {
      int planes = kRecLayers[l][1];
      std::string p = "l" + std::to_string(l) + "." + std::to_string(b);
      add(w, p + ".bn1.scale", {cin});
      add(w, p + ".bn1.shift", {cin});
      add(w, p + ".conv1.w", {planes, cin, 3, 3});
      add(w, p + ".conv1.b", {planes});
      add(w, p + ".prelu", {planes});
      add(w, p + ".conv2.w", {planes, planes, 3, 3});
      add(w, p + ".conv2.b", {planes});
      if (b == 0) {
        add(w, p + ".ds.w", {planes, cin, 1, 1});
        add(w, p + ".ds.b", {planes});
      }
      cin = planes;
    }
  add(w, "bn2.scale", {512});
  add(w, "bn2.shift", {512});
  add(w, "fc.w", {512, 512 * 7 * 7});
  add(w, "fc.b", {512});
  add(w, "feat.scale", {512});
  add(w, "feat.shift", {512});
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Filler {
  fr_weights& w;
  uint64_t seed;
  // uniform in [-a, a) + offset
  void sym(const std::string& name, float a, float offset = 0.f) {
    int ti = w.index.at(name);
    fr_tensor& t = w.tensors[ti];
    uint64_t base = splitmix64(seed ^ (0xA24BAED4963EE407ull * (uint64_t)(ti + 1)));
    const size_t n = t.data.size();
    for (size_t i = 0; i < n; ++i) {
      uint64_t h = splitmix64(base + (uint64_t)i);
      float u = (float)(h >> 40) * (1.0f / 16777216.0f);  // [0,1)
      t.data[i] = (2.f * u - 1.f) * a + offset;
    }
  }
  // uniform with target variance `var`
  void var(const std::string& name, double v) { sym(name, (float)std::sqrt(3.0 * v)); }
};

void init_det(fr_weights& w, uint64_t seed) {
  Filler f{w, seed};
  auto relu_conv = [&](const std::string& n, int fan_in) {
    f.var(n + ".w", 2.0 / fan_in);
    f.sym(n + ".b", 0.05f);
  };
  auto lin_conv = [&](const std::string& n, int fan_in) {
    f.var(n + ".w", 1.0 / fan_in);
    f.sym(n + ".b", 0.05f);
  };
  auto dwsep = [&](const std::string& n, int ci) {
    relu_conv(n + ".dw", 9);
    relu_conv(n + ".pw", ci);
  };
  // input in [-1,1]: second moment ~1/3 for uniform noise -> scale the stem up a little
  f.var("stem.w", 3.0 * 2.0 / 27.0);
  f.sym("stem.b", 0.05f);
  dwsep("b0", 16);
  int cin = 16;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kDetStages[s][0]; ++b) {
      dwsep("s" + std::to_string(s) + "." + std::to_string(b), cin);
      cin = kDetStages[s][1];
    }
  const int feats[3] = {72, 152, 288};
  for (int i = 0; i < 3; ++i) lin_conv("lat" + std::to_string(i), feats[i]);
  for (int i = 0; i < 3; ++i) lin_conv("fpn" + std::to_string(i), 16 * 9);
  for (int i = 0; i < 2; ++i) lin_conv("down" + std::to_string(i), 16 * 9);
  for (int i = 0; i < 2; ++i) lin_conv("pafpn" + std::to_string(i), 16 * 9);
  for (int i = 0; i < 3; ++i) {
    std::string h = "h" + std::to_string(i);
    dwsep(h + ".t0", 16);
    dwsep(h + ".t1", 64);
    // focal-loss style prior-probability bias (pi = 0.01) on the score logits, positive
    // mean distances so that boxes have positive extent and neighbours overlap.
    f.var(h + ".cls.w", 9.0 / (64 * 9));
    f.sym(h + ".cls.b", 0.05f, -4.595f);
    f.var(h + ".reg.w", 0.25 / (64 * 9));
    f.sym(h + ".reg.b", 0.25f, 1.5f);
    f.var(h + ".kps.w", 1.0 / (64 * 9));
    f.sym(h + ".kps.b", 0.25f, 0.f);
  }
}

void init_rec(fr_weights& w, uint64_t seed) {
  Filler f{w, seed};
  // input: uniform bytes mapped to [-1,1] -> second moment ~1/3
  f.var("stem.w", 1.0 / (27.0 / 3.0));
  f.sym("stem.b", 0.05f);
  f.sym("stem.prelu", 0.1f, 0.25f);
  double v = 0.55;  // second moment of the residual stream after the stem PReLU
  int cin = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < kRecLayers[l][0]; ++b) {
      int planes = kRecLayers[l][1];
      std::string p = "l" + std::to_string(l) + "." + std::to_string(b);
      f.sym(p + ".bn1.scale", (float)(0.2 / std::sqrt(v)), (float)(1.0 / std::sqrt(v)));
      f.sym(p + ".bn1.shift", 0.1f);
      f.var(p + ".conv1.w", 1.0 / (9.0 * cin));
      f.sym(p + ".conv1.b", 0.05f);
      f.sym(p + ".prelu", 0.1f, 0.25f);
      f.var(p + ".conv2.w", 0.5 / (9.0 * planes * 0.55));
      f.sym(p + ".conv2.b", 0.05f);
      if (b == 0) {
        f.var(p + ".ds.w", 1.0 / cin);
        f.sym(p + ".ds.b", 0.05f);
      }
      v += 0.5;
      cin = planes;
    }
  f.sym("bn2.scale", (float)(0.2 / std::sqrt(v)), (float)(1.0 / std::sqrt(v)));
  f.sym("bn2.shift", 0.1f);
  f.var("fc.w", 1.0 / 25088.0);
  f.sym("fc.b", 0.05f);
  f.sym("feat.scale", 0.2f, 1.0f);
  f.sym("feat.shift", 0.1f);
}

}  // namespace

void fr_weights_build_spec(fr_weights& w) {
  w.tensors.clear();
  w.index.clear();
  if (w.model == FR_MODEL_DET)
    build_det(w);
  else
    build_rec(w);
}

void fr_weights_random_init(fr_weights& w, uint64_t seed) {
  w.seed = seed;
  w.from_onnx = false;
  if (w.model == FR_MODEL_DET)
    init_det(w, seed);
  else
    init_rec(w, seed);
}

extern "C" {

int fr_weights_create(fr_weights** out, int model, const char* onnx_path, uint64_t seed) {
  if (!out || (model != FR_MODEL_DET && model != FR_MODEL_REC)) {
    g_werr = "fr_weights_create: bad arguments";
    return FR_ERR_INVALID_ARG;
  }
  std::unique_ptr<fr_weights> w(new fr_weights());
  w->model = model;
  fr_weights_build_spec(*w);
  if (onnx_path && onnx_path[0]) {
    std::string err;
    int s = fr_weights_load_onnx(*w, onnx_path, err);
    if (s != FR_OK) {
      g_werr = err;
      return s;
    }
    w->from_onnx = true;
  } else {
    fr_weights_random_init(*w, seed);
  }
  *out = w.release();
  return FR_OK;
}

void fr_weights_destroy(fr_weights* w) { delete w; }
int fr_weights_model(const fr_weights* w) { return w ? w->model : -1; }
int fr_weights_from_onnx(const fr_weights* w) { return w && w->from_onnx ? 1 : 0; }
int fr_weights_num_tensors(const fr_weights* w) { return w ? (int)w->tensors.size() : 0; }

int fr_weights_tensor_info(const fr_weights* w, int idx, char* name, int name_cap, int64_t dims[4],
                           int* ndim) {
  if (!w || idx < 0 || idx >= (int)w->tensors.size()) return FR_ERR_INVALID_ARG;
  const fr_tensor& t = w->tensors[idx];
  if (name && name_cap > 0) {
    strncpy(name, t.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (dims)
    for (int i = 0; i < 4; ++i) dims[i] = i < (int)t.dims.size() ? t.dims[i] : 1;
  if (ndim) *ndim = (int)t.dims.size();
  return FR_OK;
}

int fr_weights_tensor_get(const fr_weights* w, int idx, float* out, size_t n) {
  if (!w || !out || idx < 0 || idx >= (int)w->tensors.size()) return FR_ERR_INVALID_ARG;
  const fr_tensor& t = w->tensors[idx];
  if (n != t.data.size()) return FR_ERR_CAPACITY;
  memcpy(out, t.data.data(), n * sizeof(float));
  return FR_OK;
}

int fr_weights_tensor_set(fr_weights* w, int idx, const float* data, size_t n) {
  if (!w || !data || idx < 0 || idx >= (int)w->tensors.size()) return FR_ERR_INVALID_ARG;
  fr_tensor& t = w->tensors[idx];
  if (n != t.data.size()) return FR_ERR_CAPACITY;
  memcpy(t.data.data(), data, n * sizeof(float));
  return FR_OK;
}

const char* fr_weights_last_error(void) { return g_werr.c_str(); }

}  // extern "C"
