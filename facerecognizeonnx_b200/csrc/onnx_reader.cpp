// ONNX initializer ingestion without protobuf / onnx / ONNX Runtime: a minimal wire-format
// reader for the fields we need, replacing the model-file half of
// FaceDetector::loadModel / FaceRecognizer::loadModel (reference src/face_detector.cpp:20-90,
// src/face_recognizer.cpp:21-91: open the file, introspect input/outputs).
//
// The two graphs are fixed architectures, so tensors are matched to the canonical list
// (weights.cpp / oracle/nets.py) by walking the nodes in graph order and checking every shape:
//   rec (arcface iresnet50 export): Conv(+bias, BN folded) / PRelu / BatchNormalization / Gemm
//   det (scrfd_500m_bnkps export) : Conv(+bias, BN folded) in execution order, optional scalar
//                                   Mul after each bbox conv (mmdet Scale layer) folded in
// Anything unexpected fails loudly with FR_ERR_MODEL (loadModel -> false), never silently.
#include <cmath>
#include <cstring>
#include <fstream>

#include "common.h"

namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool eof() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end && shift < 64) {
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
    ok = false;
    return 0;
  }
  Reader sub() {
    const uint64_t n = varint();
    if (!ok || n > (uint64_t)(end - p)) {
      ok = false;
      return Reader{end, end, false};
    }
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wire) {
    if (wire == 0) varint();
    else if (wire == 1) p += 8;
    else if (wire == 2) sub();
    else if (wire == 5) p += 4;
    else ok = false;
    if (p > end) ok = false;
  }
};

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int data_type = 0;
  std::vector<float> data;
};

struct OnnxNode {
  std::string op;
  std::vector<std::string> inputs;
  float epsilon = 1e-5f;
  int64_t transB = 0;
  int64_t group = 1;
};

std::string str(Reader r) { return std::string(reinterpret_cast<const char*>(r.p), r.end - r.p); }

bool parse_tensor(Reader r, OnnxTensor& t) {
  std::vector<float> fdata;
  Reader raw{nullptr, nullptr};
  bool have_raw = false;
  while (!r.eof() && r.ok) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1 && wire == 0) t.dims.push_back((int64_t)r.varint());
    else if (field == 1 && wire == 2) { Reader d = r.sub(); while (!d.eof() && d.ok) t.dims.push_back((int64_t)d.varint()); }
    else if (field == 2 && wire == 0) t.data_type = (int)r.varint();
    else if (field == 4 && wire == 2) { Reader d = r.sub(); while (d.p + 4 <= d.end) { float f; memcpy(&f, d.p, 4); fdata.push_back(f); d.p += 4; } }
    else if (field == 4 && wire == 5) { float f; memcpy(&f, r.p, 4); r.p += 4; fdata.push_back(f); }
    else if (field == 8 && wire == 2) t.name = str(r.sub());
    else if (field == 9 && wire == 2) { raw = r.sub(); have_raw = true; }
    else r.skip(wire);
  }
  if (!r.ok) return false;
  if (t.data_type != 1) return true;  // only float tensors carry weights we need
  size_t n = 1;
  for (auto d : t.dims) n *= (size_t)d;
  if (have_raw) {
    if ((size_t)(raw.end - raw.p) != n * 4) return false;
    t.data.resize(n);
    memcpy(t.data.data(), raw.p, n * 4);
  } else {
    if (fdata.size() != n) return false;
    t.data = std::move(fdata);
  }
  return true;
}

bool parse_node(Reader r, OnnxNode& n) {
  while (!r.eof() && r.ok) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1 && wire == 2) n.inputs.push_back(str(r.sub()));
    else if (field == 4 && wire == 2) n.op = str(r.sub());
    else if (field == 5 && wire == 2) {
      Reader a = r.sub();
      std::string aname;
      float f = 0;
      int64_t i = 0;
      bool hf = false, hi = false;
      while (!a.eof() && a.ok) {
        const uint64_t k2 = a.varint();
        const int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
        if (f2 == 1 && w2 == 2) aname = str(a.sub());
        else if (f2 == 2 && w2 == 5) { memcpy(&f, a.p, 4); a.p += 4; hf = true; }
        else if (f2 == 3 && w2 == 0) { i = (int64_t)a.varint(); hi = true; }
        else a.skip(w2);
      }
      if (aname == "epsilon" && hf) n.epsilon = f;
      if (aname == "transB" && hi) n.transB = i;
      if (aname == "group" && hi) n.group = i;
    } else r.skip(wire);
  }
  return r.ok;
}

struct Graph {
  std::vector<OnnxNode> nodes;
  std::map<std::string, OnnxTensor> init;
};

bool parse_model(const std::vector<uint8_t>& buf, Graph& g, std::string& err) {
  Reader m{buf.data(), buf.data() + buf.size()};
  bool found = false;
  while (!m.eof() && m.ok) {
    const uint64_t key = m.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 7 && wire == 2) {
      found = true;
      Reader gr = m.sub();
      while (!gr.eof() && gr.ok) {
        const uint64_t k2 = gr.varint();
        const int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
        if (f2 == 1 && w2 == 2) {
          OnnxNode n;
          if (!parse_node(gr.sub(), n)) { err = "malformed NodeProto"; return false; }
          g.nodes.push_back(std::move(n));
        } else if (f2 == 5 && w2 == 2) {
          OnnxTensor t;
          if (!parse_tensor(gr.sub(), t)) { err = "malformed TensorProto " + t.name; return false; }
          g.init[t.name] = std::move(t);
        } else gr.skip(w2);
      }
      if (!gr.ok) { err = "malformed GraphProto"; return false; }
    } else m.skip(wire);
  }
  if (!m.ok || !found) { err = "not an ONNX ModelProto (no graph)"; return false; }
  return true;
}

const OnnxTensor* init_of(const Graph& g, const OnnxNode& n, size_t idx) {
  if (idx >= n.inputs.size()) return nullptr;
  auto it = g.init.find(n.inputs[idx]);
  return it == g.init.end() || it->second.data.empty() ? nullptr : &it->second;
}

bool assign(fr_weights& w, const std::string& name, const std::vector<float>& data, std::string& err) {
  auto it = w.index.find(name);
  if (it == w.index.end()) { err = "internal: unknown tensor " + name; return false; }
  fr_tensor& t = w.tensors[it->second];
  if (t.data.size() != data.size()) {
    err = "shape mismatch for " + name + ": file has " + std::to_string(data.size()) + " elements, architecture needs " +
          std::to_string(t.data.size());
    return false;
  }
  t.data = data;
  return true;
}

bool same_dims(const OnnxTensor& t, const fr_tensor& want) {
  size_t a = 1, b = 1;
  for (auto d : t.dims) a *= (size_t)d;
  for (auto d : want.dims) b *= (size_t)d;
  if (a != b) return false;
  // conv weights must agree dimension by dimension; vectors may come as [C], [C,1,1] ...
  if (want.dims.size() == 4) return t.dims.size() == 4 && std::equal(t.dims.begin(), t.dims.end(), want.dims.begin());
  return true;
}

// Conv weight + bias (zero bias if the node has none) into `<name>.w` / `<name>.b`
bool take_conv(const Graph& g, const OnnxNode& n, fr_weights& w, const std::string& name, std::string& err) {
  const OnnxTensor* wt = init_of(g, n, 1);
  if (!wt) { err = "Conv for " + name + " has no weight initializer"; return false; }
  const fr_tensor& want = w.at(name + ".w");
  if (!same_dims(*wt, want)) {
    err = "Conv " + name + ": unexpected weight shape";
    return false;
  }
  if (!assign(w, name + ".w", wt->data, err)) return false;
  const OnnxTensor* bt = init_of(g, n, 2);
  std::vector<float> b = bt ? bt->data : std::vector<float>((size_t)want.dims[0], 0.f);
  return assign(w, name + ".b", b, err);
}

bool take_bn(const Graph& g, const OnnxNode& n, fr_weights& w, const std::string& name, std::string& err) {
  const OnnxTensor *sc = init_of(g, n, 1), *bi = init_of(g, n, 2), *mu = init_of(g, n, 3), *var = init_of(g, n, 4);
  if (!sc || !bi || !mu || !var) { err = "BatchNormalization for " + name + " lacks initializers"; return false; }
  const size_t c = sc->data.size();
  if (bi->data.size() != c || mu->data.size() != c || var->data.size() != c) { err = "BatchNormalization " + name + ": ragged"; return false; }
  std::vector<float> s(c), t(c);
  for (size_t i = 0; i < c; ++i) {
    s[i] = sc->data[i] / std::sqrt(var->data[i] + n.epsilon);
    t[i] = bi->data[i] - mu->data[i] * s[i];
  }
  return assign(w, name + ".scale", s, err) && assign(w, name + ".shift", t, err);
}

int load_rec(const Graph& g, fr_weights& w, std::string& err) {
  std::vector<const OnnxNode*> conv3, conv1, bn, prelu, gemm;
  for (const OnnxNode& n : g.nodes) {
    if (n.op == "Conv") {
      const OnnxTensor* wt = init_of(g, n, 1);
      if (!wt || wt->dims.size() != 4) { err = "Conv without 4-D weight initializer"; return FR_ERR_MODEL; }
      (wt->dims[2] == 1 ? conv1 : conv3).push_back(&n);
    } else if (n.op == "BatchNormalization") bn.push_back(&n);
    else if (n.op == "PRelu") prelu.push_back(&n);
    else if (n.op == "Gemm") gemm.push_back(&n);
  }
  if (conv3.size() != 49 || conv1.size() != 4 || bn.size() != 26 || prelu.size() != 25 || gemm.size() != 1) {
    err = "not the expected IResNet-50 export: " + std::to_string(conv3.size()) + " 3x3 convs (want 49), " +
          std::to_string(conv1.size()) + " 1x1 convs (want 4), " + std::to_string(bn.size()) +
          " BatchNormalization (want 26), " + std::to_string(prelu.size()) + " PRelu (want 25), " +
          std::to_string(gemm.size()) + " Gemm (want 1)";
    return FR_ERR_MODEL;
  }
  auto take_prelu = [&](const OnnxNode& n, const std::string& name) {
    const OnnxTensor* s = init_of(g, n, 1);
    if (!s) { err = "PRelu for " + name + " has no slope initializer"; return false; }
    return assign(w, name, s->data, err);
  };
  // stem conv is stored as "stem.w"/"stem.b"
  {
    const OnnxTensor* wt = init_of(g, *conv3[0], 1);
    const OnnxTensor* bt = init_of(g, *conv3[0], 2);
    if (!wt || !same_dims(*wt, w.at("stem.w"))) { err = "stem conv: unexpected weight shape"; return FR_ERR_MODEL; }
    if (!assign(w, "stem.w", wt->data, err)) return FR_ERR_MODEL;
    if (!assign(w, "stem.b", bt ? bt->data : std::vector<float>(64, 0.f), err)) return FR_ERR_MODEL;
    if (!take_prelu(*prelu[0], "stem.prelu")) return FR_ERR_MODEL;
  }
  const int layers[4] = {3, 4, 14, 3};
  int bi = 0, ci = 1, pi = 1, di = 0;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < layers[l]; ++b, ++bi) {
      const std::string p = "l" + std::to_string(l) + "." + std::to_string(b);
      if (!take_bn(g, *bn[bi], w, p + ".bn1", err)) return FR_ERR_MODEL;
      if (!take_conv(g, *conv3[ci++], w, p + ".conv1", err)) return FR_ERR_MODEL;
      if (!take_prelu(*prelu[pi++], p + ".prelu")) return FR_ERR_MODEL;
      if (!take_conv(g, *conv3[ci++], w, p + ".conv2", err)) return FR_ERR_MODEL;
      if (b == 0 && !take_conv(g, *conv1[di++], w, p + ".ds", err)) return FR_ERR_MODEL;
    }
  if (!take_bn(g, *bn[24], w, "bn2", err)) return FR_ERR_MODEL;
  {
    const OnnxNode& n = *gemm[0];
    const OnnxTensor *wt = init_of(g, n, 1), *bt = init_of(g, n, 2);
    if (!wt || wt->dims.size() != 2) { err = "Gemm without 2-D weight"; return FR_ERR_MODEL; }
    std::vector<float> fw;
    if (n.transB) {
      if (wt->dims[0] != 512 || wt->dims[1] != 25088) { err = "Gemm weight shape is not [512,25088]"; return FR_ERR_MODEL; }
      fw = wt->data;
    } else {
      if (wt->dims[0] != 25088 || wt->dims[1] != 512) { err = "Gemm weight shape is not [25088,512]"; return FR_ERR_MODEL; }
      fw.resize(wt->data.size());
      for (int k = 0; k < 25088; ++k)
        for (int o = 0; o < 512; ++o) fw[(size_t)o * 25088 + k] = wt->data[(size_t)k * 512 + o];
    }
    if (!assign(w, "fc.w", fw, err)) return FR_ERR_MODEL;
    if (!assign(w, "fc.b", bt ? bt->data : std::vector<float>(512, 0.f), err)) return FR_ERR_MODEL;
  }
  {
    // final BatchNormalization -> feat.scale / feat.shift
    fr_weights tmp;  // reuse take_bn through a name shim
    const OnnxNode& n = *bn[25];
    const OnnxTensor *sc = init_of(g, n, 1), *bb = init_of(g, n, 2), *mu = init_of(g, n, 3), *var = init_of(g, n, 4);
    if (!sc || !bb || !mu || !var || sc->data.size() != 512) { err = "features BatchNormalization malformed"; return FR_ERR_MODEL; }
    std::vector<float> s(512), t(512);
    for (int i = 0; i < 512; ++i) {
      s[i] = sc->data[i] / std::sqrt(var->data[i] + n.epsilon);
      t[i] = bb->data[i] - mu->data[i] * s[i];
    }
    if (!assign(w, "feat.scale", s, err) || !assign(w, "feat.shift", t, err)) return FR_ERR_MODEL;
  }
  return FR_OK;
}

int load_det(const Graph& g, fr_weights& w, std::string& err) {
  // Conv nodes in execution order map 1:1 onto the canonical ".w" tensors in list order.
  std::vector<std::string> names;
  for (const fr_tensor& t : w.tensors)
    if (t.name.size() > 2 && t.name.compare(t.name.size() - 2, 2, ".w") == 0) names.push_back(t.name.substr(0, t.name.size() - 2));
  std::vector<const OnnxNode*> convs;
  std::vector<float> scales;  // scalar Mul initializers in graph order (mmdet Scale on bbox_pred)
  for (const OnnxNode& n : g.nodes) {
    if (n.op == "Conv") convs.push_back(&n);
    else if (n.op == "Mul")
      for (size_t i = 0; i < n.inputs.size(); ++i) {
        const OnnxTensor* t = init_of(g, n, i);
        if (t && t->data.size() == 1) scales.push_back(t->data[0]);
      }
  }
  if (convs.size() != names.size()) {
    err = "not the expected SCRFD-500M export: " + std::to_string(convs.size()) + " Conv nodes, architecture has " +
          std::to_string(names.size());
    return FR_ERR_MODEL;
  }
  for (size_t i = 0; i < names.size(); ++i)
    if (!take_conv(g, *convs[i], w, names[i], err)) return FR_ERR_MODEL;
  if (scales.size() == 3) {
    for (int s = 0; s < 3; ++s) {
      const std::string n = "h" + std::to_string(s) + ".reg";
      for (float& v : w.tensors[w.index.at(n + ".w")].data) v *= scales[s];
      for (float& v : w.tensors[w.index.at(n + ".b")].data) v *= scales[s];
    }
  } else if (!scales.empty()) {
    err = "unexpected number of scalar Mul nodes (" + std::to_string(scales.size()) + "; want 0 or 3)";
    return FR_ERR_MODEL;
  }
  return FR_OK;
}

}  // namespace

int fr_weights_load_onnx(fr_weights& w, const char* path, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    err = std::string("cannot open ") + path;
    return FR_ERR_MODEL;
  }
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  Graph g;
  if (!parse_model(buf, g, err)) {
    err = std::string(path) + ": " + err;
    return FR_ERR_MODEL;
  }
  const int s = w.model == FR_MODEL_DET ? load_det(g, w, err) : load_rec(g, w, err);
  if (s != FR_OK) err = std::string(path) + ": " + err;
  return s;
}
