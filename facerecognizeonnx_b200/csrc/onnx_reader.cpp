// ONNX initializer ingestion without protobuf / onnx / ONNX Runtime: a minimal wire-format
// reader for the fields we need, replacing the model-file half of
// FaceDetector::loadModel / FaceRecognizer::loadModel (reference src/face_detector.cpp:20-90,
// src/face_recognizer.cpp:21-91: open the file, introspect input/outputs).
//
// The two graphs are fixed architectures.  Tensors are matched to the canonical list
// (weights.cpp) by FOLLOWING THE GRAPH'S EDGES from its input -- never by the order in which
// nodes happen to be stored -- and every weight shape / stride / group is checked on the way:
//   rec (arcface iresnet50): Conv [-> BatchNormalization] -> PRelu -> { BatchNormalization ->
//        Conv [-> BN] -> PRelu -> Conv [-> BN] ; shortcut = identity | Conv 1x1 [-> BN] ; Add } x 24
//        -> BatchNormalization -> Flatten | Reshape -> Gemm -> BatchNormalization
//   det (scrfd_500m_bnkps): Conv [-> BN] -> Relu backbone of depthwise-separable blocks, PAFPN
//        (1x1 laterals, Resize + Add top-down, 3x3 fpn convs, stride-2 3x3 + Add bottom-up, pafpn
//        convs), per-stride towers and 3x3 predictors, optional scalar Mul (mmdet Scale) after bbox.
// A BatchNormalization that directly follows a Conv is folded into it here (double precision), so
// both raw training-graph exports and exporter-fused files load to the same canonical tensors.
// Anything unexpected fails loudly with FR_ERR_MODEL (loadModel -> false), never silently.
#include <cmath>
#include <cstring>
#include <fstream>
#include <set>

#include "common.h"

namespace {

// ------------------------------------------------------------------ wire --
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool eof() const { return !ok || p >= end; }
  size_t left() const { return (size_t)(end - p); }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (ok && p < end && shift < 64) {
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
    if (!ok || n > (uint64_t)left()) {
      ok = false;
      return Reader{end, end, false};
    }
    Reader r{p, p + n};
    p += n;
    return r;
  }
  bool fixed32(void* out) {
    if (!ok || left() < 4) {
      ok = false;
      return false;
    }
    memcpy(out, p, 4);
    p += 4;
    return true;
  }
  void advance(size_t n) {
    if (!ok || left() < n) ok = false;
    else p += n;
  }
  void skip(int wire) {
    if (wire == 0) varint();
    else if (wire == 1) advance(8);
    else if (wire == 2) sub();
    else if (wire == 5) advance(4);
    else ok = false;
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
  std::vector<std::string> inputs, outputs;
  float epsilon = 1e-5f;
  int64_t transB = 0;
  int64_t group = 1;
  std::vector<int64_t> strides;
  int stride() const { return strides.empty() ? 1 : (int)strides[0]; }
};

std::string str(Reader r) { return r.ok ? std::string(reinterpret_cast<const char*>(r.p), r.left()) : std::string(); }

constexpr size_t kMaxElems = (size_t)1 << 28;  // 1 GiB of floats: far above fc.w (12.8 M)

bool parse_tensor(Reader r, OnnxTensor& t) {
  std::vector<float> fdata;
  Reader raw{nullptr, nullptr};
  bool have_raw = false;
  while (!r.eof()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1 && wire == 0) t.dims.push_back((int64_t)r.varint());
    else if (field == 1 && wire == 2) { Reader d = r.sub(); while (!d.eof()) t.dims.push_back((int64_t)d.varint()); if (!d.ok) r.ok = false; }
    else if (field == 2 && wire == 0) t.data_type = (int)r.varint();
    else if (field == 4 && wire == 2) { Reader d = r.sub(); float f; while (d.ok && d.left() >= 4 && d.fixed32(&f)) fdata.push_back(f); if (!d.ok || d.left()) r.ok = false; }
    else if (field == 4 && wire == 5) { float f; if (r.fixed32(&f)) fdata.push_back(f); }
    else if (field == 8 && wire == 2) t.name = str(r.sub());
    else if (field == 9 && wire == 2) { raw = r.sub(); have_raw = true; }
    else r.skip(wire);
  }
  if (!r.ok) return false;
  if (t.data_type != 1) return true;  // only float tensors carry weights we need
  size_t n = 1;
  for (auto d : t.dims) {
    if (d < 0 || (uint64_t)d > kMaxElems) return false;
    n *= (size_t)d;
    if (n > kMaxElems) return false;
  }
  if (have_raw) {
    if (!raw.ok || raw.left() != n * 4) return false;
    t.data.resize(n);
    if (n) memcpy(t.data.data(), raw.p, n * 4);
  } else {
    if (fdata.size() != n) return false;
    t.data = std::move(fdata);
  }
  return true;
}

bool parse_node(Reader r, OnnxNode& n) {
  while (!r.eof()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1 && wire == 2) n.inputs.push_back(str(r.sub()));
    else if (field == 2 && wire == 2) n.outputs.push_back(str(r.sub()));
    else if (field == 4 && wire == 2) n.op = str(r.sub());
    else if (field == 5 && wire == 2) {
      Reader a = r.sub();
      std::string aname;
      float f = 0;
      int64_t i = 0;
      std::vector<int64_t> ints;
      bool hf = false, hi = false;
      while (!a.eof()) {
        const uint64_t k2 = a.varint();
        const int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
        if (f2 == 1 && w2 == 2) aname = str(a.sub());
        else if (f2 == 2 && w2 == 5) hf = a.fixed32(&f);
        else if (f2 == 3 && w2 == 0) { i = (int64_t)a.varint(); hi = true; }
        else if (f2 == 8 && w2 == 0) ints.push_back((int64_t)a.varint());
        else if (f2 == 8 && w2 == 2) { Reader d = a.sub(); while (!d.eof()) ints.push_back((int64_t)d.varint()); if (!d.ok) a.ok = false; }
        else a.skip(w2);
      }
      if (!a.ok) r.ok = false;
      if (aname == "epsilon" && hf) n.epsilon = f;
      if (aname == "transB" && hi) n.transB = i;
      if (aname == "group" && hi) n.group = i;
      if (aname == "strides") n.strides = ints;
    } else r.skip(wire);
  }
  return r.ok;
}

struct Graph {
  std::vector<OnnxNode> nodes;
  std::map<std::string, OnnxTensor> init;
  std::map<std::string, int> producer;                 // tensor -> node
  std::map<std::string, std::vector<int>> consumers;   // tensor -> nodes reading it as a data input
  std::string input;                                    // the one tensor nobody produces
};

bool parse_model(const std::vector<uint8_t>& buf, Graph& g, std::string& err) {
  Reader m{buf.data(), buf.data() + buf.size()};
  bool found = false;
  while (!m.eof()) {
    const uint64_t key = m.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 7 && wire == 2) {
      found = true;
      Reader gr = m.sub();
      while (!gr.eof()) {
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
  for (size_t i = 0; i < g.nodes.size(); ++i)
    for (const std::string& o : g.nodes[i].outputs) g.producer[o] = (int)i;
  std::set<std::string> free_inputs;
  for (size_t i = 0; i < g.nodes.size(); ++i)
    for (const std::string& in : g.nodes[i].inputs) {
      if (in.empty() || g.init.count(in)) continue;
      g.consumers[in].push_back((int)i);
      if (!g.producer.count(in)) free_inputs.insert(in);
    }
  if (free_inputs.size() != 1) {
    err = "expected exactly one graph input, found " + std::to_string(free_inputs.size());
    return false;
  }
  g.input = *free_inputs.begin();
  return true;
}

// ------------------------------------------------------------------ matcher --
struct Matcher {
  const Graph& g;
  fr_weights& w;
  std::string err;
  std::set<int> used;  // nodes already bound to a canonical layer

  Matcher(const Graph& g_, fr_weights& w_) : g(g_), w(w_) {}
  bool fail(const std::string& m) {
    if (err.empty()) err = m;
    return false;
  }
  const OnnxTensor* init_of(const OnnxNode& n, size_t idx) const {
    if (idx >= n.inputs.size()) return nullptr;
    auto it = g.init.find(n.inputs[idx]);
    return it == g.init.end() || it->second.data.empty() ? nullptr : &it->second;
  }
  const std::vector<int>& cons(const std::string& t) const {
    static const std::vector<int> none;
    auto it = g.consumers.find(t);
    return it == g.consumers.end() ? none : it->second;
  }
  // the single unused consumer of `t` with op `op` (and, for Conv, the given weight dims /
  // stride / group); -1 if none, -2 if ambiguous
  int find(const std::string& t, const std::string& op, const std::vector<int64_t>* wdims = nullptr, int stride = 0,
           int group = 0) const {
    int hit = -1;
    for (int ni : cons(t)) {
      const OnnxNode& n = g.nodes[ni];
      if (n.op != op || used.count(ni)) continue;
      if (wdims) {
        auto it = n.inputs.size() > 1 ? g.init.find(n.inputs[1]) : g.init.end();
        if (it == g.init.end() || it->second.dims != *wdims) continue;
      }
      if (stride && n.stride() != stride) continue;
      if (group && n.group != group) continue;
      if (hit != -1) return -2;
      hit = ni;
    }
    return hit;
  }
  bool assign(const std::string& name, const std::vector<float>& data) {
    auto it = w.index.find(name);
    if (it == w.index.end()) return fail("internal: unknown tensor " + name);
    fr_tensor& t = w.tensors[it->second];
    if (t.data.size() != data.size())
      return fail("shape mismatch for " + name + ": file has " + std::to_string(data.size()) +
                  " elements, architecture needs " + std::to_string(t.data.size()));
    t.data = data;
    return true;
  }
  bool bn_affine(const OnnxNode& n, const std::string& what, std::vector<double>& s, std::vector<double>& t) {
    const OnnxTensor *sc = init_of(n, 1), *bi = init_of(n, 2), *mu = init_of(n, 3), *var = init_of(n, 4);
    if (!sc || !bi || !mu || !var) return fail("BatchNormalization for " + what + " lacks initializers");
    const size_t c = sc->data.size();
    if (bi->data.size() != c || mu->data.size() != c || var->data.size() != c)
      return fail("BatchNormalization " + what + ": ragged parameters");
    s.resize(c);
    t.resize(c);
    for (size_t i = 0; i < c; ++i) {
      s[i] = (double)sc->data[i] / std::sqrt((double)var->data[i] + (double)n.epsilon);
      t[i] = (double)bi->data[i] - (double)mu->data[i] * s[i];
    }
    return true;
  }
  // x -> BatchNormalization -> out; stores <name>.scale / <name>.shift
  bool take_bn(const std::string& x, const std::string& name, std::string& out) {
    const int ni = find(x, "BatchNormalization");
    if (ni < 0) return fail("expected one BatchNormalization (" + name + ") after tensor '" + x + "'");
    std::vector<double> s, t;
    if (!bn_affine(g.nodes[ni], name, s, t)) return false;
    std::vector<float> sf(s.begin(), s.end()), tf(t.begin(), t.end());
    used.insert(ni);
    out = g.nodes[ni].outputs.empty() ? std::string() : g.nodes[ni].outputs[0];
    return assign(name + ".scale", sf) && assign(name + ".shift", tf);
  }
  // x -> Conv(<name>) [-> BatchNormalization folded] -> out
  bool take_conv(const std::string& x, const std::string& name, int stride, std::string& out) {
    const fr_tensor& want = w.at(name + ".w");
    const int group = want.dims[1] == 1 && want.dims[0] > 1 && want.dims[2] == 3 ? (int)want.dims[0] : 1;
    const int ni = find(x, "Conv", &want.dims, stride, group);
    if (ni == -2) return fail("ambiguous Conv candidates for " + name + " after tensor '" + x + "'");
    if (ni < 0) {
      std::string have;
      for (int ci : cons(x)) {
        const OnnxNode& n = g.nodes[ci];
        have += " " + n.op;
        auto it = n.op == "Conv" && n.inputs.size() > 1 ? g.init.find(n.inputs[1]) : g.init.end();
        if (it != g.init.end()) {
          have += "[";
          for (auto d : it->second.dims) have += std::to_string(d) + ",";
          have += "s" + std::to_string(n.stride()) + " g" + std::to_string(n.group) + "]";
        }
      }
      std::string dims;
      for (auto d : want.dims) dims += std::to_string(d) + ",";
      return fail("no Conv matching " + name + " (weight [" + dims + "] stride " + std::to_string(stride) + " group " +
                  std::to_string(group) + ") reads tensor '" + x + "'; consumers:" + (have.empty() ? " none" : have));
    }
    const OnnxNode& n = g.nodes[ni];
    used.insert(ni);
    const OnnxTensor* wt = init_of(n, 1);
    if (!wt) return fail("Conv for " + name + " has no float weight initializer");
    const OnnxTensor* bt = init_of(n, 2);
    const size_t co = (size_t)want.dims[0], per = wt->data.size() / co;
    if (bt && bt->data.size() != co) return fail("Conv " + name + ": bias size");
    std::vector<float> wv = wt->data;
    std::vector<float> bv = bt ? bt->data : std::vector<float>(co, 0.f);
    out = n.outputs.empty() ? std::string() : n.outputs[0];
    // fold a directly following BatchNormalization (the only consumer of the conv output)
    const std::vector<int>& cs = cons(out);
    if (cs.size() == 1 && g.nodes[cs[0]].op == "BatchNormalization" && !used.count(cs[0])) {
      const OnnxNode& b = g.nodes[cs[0]];
      std::vector<double> s, t;
      if (!bn_affine(b, name, s, t)) return false;
      if (s.size() != co) return fail("BatchNormalization after " + name + ": channel count");
      for (size_t o = 0; o < co; ++o) {
        for (size_t k = 0; k < per; ++k) wv[o * per + k] = (float)((double)wv[o * per + k] * s[o]);
        bv[o] = (float)((double)bv[o] * s[o] + t[o]);
      }
      used.insert(cs[0]);
      out = b.outputs.empty() ? std::string() : b.outputs[0];
    }
    return assign(name + ".w", wv) && assign(name + ".b", bv);
  }
  // x -> op -> out (a parameter-free node such as Relu / Add / Flatten)
  bool through(const std::string& x, const std::string& op, std::string& out, const char* alt = nullptr) {
    int ni = find(x, op);
    if (ni < 0 && alt) ni = find(x, alt);
    if (ni < 0) return fail("expected " + op + " after tensor '" + x + "'");
    used.insert(ni);
    out = g.nodes[ni].outputs.empty() ? std::string() : g.nodes[ni].outputs[0];
    return true;
  }
  bool conv_relu(const std::string& x, const std::string& name, int stride, std::string& out) {
    std::string y;
    return take_conv(x, name, stride, y) && through(y, "Relu", out);
  }
  bool dwsep(const std::string& x, const std::string& name, int stride, std::string& out) {
    std::string y;
    return conv_relu(x, name + ".dw", stride, y) && conv_relu(y, name + ".pw", 1, out);
  }
  bool take_prelu(const std::string& x, const std::string& name, std::string& out) {
    const int ni = find(x, "PRelu");
    if (ni < 0) return fail("expected PRelu (" + name + ") after tensor '" + x + "'");
    const OnnxTensor* s = init_of(g.nodes[ni], 1);
    if (!s) return fail("PRelu for " + name + " has no slope initializer");
    used.insert(ni);
    out = g.nodes[ni].outputs[0];
    return assign(name, s->data);
  }
  // the Add that joins a and b
  bool join(const std::string& a, const std::string& b, std::string& out) {
    for (int ni : cons(a)) {
      const OnnxNode& n = g.nodes[ni];
      if (n.op != "Add" || used.count(ni) || n.inputs.size() != 2) continue;
      if ((n.inputs[0] == a && n.inputs[1] == b) || (n.inputs[0] == b && n.inputs[1] == a)) {
        used.insert(ni);
        out = n.outputs[0];
        return true;
      }
    }
    return fail("expected Add(" + a + ", " + b + ")");
  }
};

int load_rec(const Graph& g, fr_weights& w, std::string& err) {
  Matcher m(g, w);
  auto bad = [&]() {
    err = "not the expected IResNet-50 export: " + m.err;
    return FR_ERR_MODEL;
  };
  std::string x, y, z;
  if (!m.take_conv(g.input, "stem", 1, y) || !m.take_prelu(y, "stem.prelu", x)) return bad();
  const int layers[4] = {3, 4, 14, 3};
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < layers[l]; ++b) {
      const std::string p = "l" + std::to_string(l) + "." + std::to_string(b);
      std::string sc = x;
      if (!m.take_bn(x, p + ".bn1", y) || !m.take_conv(y, p + ".conv1", 1, z) || !m.take_prelu(z, p + ".prelu", y) ||
          !m.take_conv(y, p + ".conv2", b == 0 ? 2 : 1, z))
        return bad();
      if (b == 0 && !m.take_conv(x, p + ".ds", 2, sc)) return bad();
      if (!m.join(z, sc, x)) return bad();
    }
  if (!m.take_bn(x, "bn2", y) || !m.through(y, "Flatten", z, "Reshape")) return bad();
  {
    const int ni = m.find(z, "Gemm");
    if (ni < 0) { m.fail("expected Gemm after the flattened features"); return bad(); }
    const OnnxNode& n = g.nodes[ni];
    const OnnxTensor *wt = m.init_of(n, 1), *bt = m.init_of(n, 2);
    if (!wt || wt->dims.size() != 2) { m.fail("Gemm without 2-D weight"); return bad(); }
    std::vector<float> fw;
    if (n.transB) {
      if (wt->dims[0] != 512 || wt->dims[1] != 25088) { m.fail("Gemm weight shape is not [512,25088]"); return bad(); }
      fw = wt->data;
    } else {
      if (wt->dims[0] != 25088 || wt->dims[1] != 512) { m.fail("Gemm weight shape is not [25088,512]"); return bad(); }
      fw.resize(wt->data.size());
      for (int k = 0; k < 25088; ++k)
        for (int o = 0; o < 512; ++o) fw[(size_t)o * 25088 + k] = wt->data[(size_t)k * 512 + o];
    }
    if (!m.assign("fc.w", fw) || !m.assign("fc.b", bt ? bt->data : std::vector<float>(512, 0.f))) return bad();
    x = n.outputs.empty() ? std::string() : n.outputs[0];
  }
  if (!m.take_bn(x, "feat", y)) return bad();
  return FR_OK;
}

int load_det(const Graph& g, fr_weights& w, std::string& err) {
  Matcher m(g, w);
  auto bad = [&]() {
    err = "not the expected SCRFD-500M export: " + m.err;
    return FR_ERR_MODEL;
  };
  const int stages[4] = {2, 3, 2, 6};
  std::string x, y;
  if (!m.conv_relu(g.input, "stem", 2, x) || !m.dwsep(x, "b0", 1, y)) return bad();
  x = y;
  std::string feats[3], lat[3], inter[3], outs[3];
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < stages[s]; ++b) {
      if (!m.dwsep(x, "s" + std::to_string(s) + "." + std::to_string(b), b == 0 ? 2 : 1, y)) return bad();
      x = y;
    }
    if (s >= 1) feats[s - 1] = x;
  }
  for (int i = 0; i < 3; ++i)
    if (!m.take_conv(feats[i], "lat" + std::to_string(i), 1, lat[i])) return bad();
  // top-down: merged[i-1] = lat[i-1] + upsample(merged[i])
  for (int i = 2; i >= 1; --i) {
    std::string up;
    if (!m.through(lat[i], "Resize", up, "Upsample") || !m.join(lat[i - 1], up, y)) return bad();
    lat[i - 1] = y;
  }
  for (int i = 0; i < 3; ++i)
    if (!m.take_conv(lat[i], "fpn" + std::to_string(i), 1, inter[i])) return bad();
  // bottom-up: inter[i+1] += down_i(inter[i])
  for (int i = 0; i < 2; ++i) {
    std::string d;
    if (!m.take_conv(inter[i], "down" + std::to_string(i), 2, d) || !m.join(inter[i + 1], d, y)) return bad();
    inter[i + 1] = y;
  }
  outs[0] = inter[0];
  for (int i = 1; i < 3; ++i)
    if (!m.take_conv(inter[i], "pafpn" + std::to_string(i - 1), 1, outs[i])) return bad();
  for (int i = 0; i < 3; ++i) {
    const std::string h = "h" + std::to_string(i);
    std::string t;
    if (!m.dwsep(outs[i], h + ".t0", 1, y) || !m.dwsep(y, h + ".t1", 1, t)) return bad();
    std::string o;
    if (!m.take_conv(t, h + ".cls", 1, o) || !m.take_conv(t, h + ".kps", 1, o) || !m.take_conv(t, h + ".reg", 1, o))
      return bad();
    // mmdet Scale layer: bbox_pred * scalar
    const int mi = m.find(o, "Mul");
    if (mi >= 0) {
      const OnnxNode& mul = g.nodes[mi];
      const OnnxTensor* sc = nullptr;
      for (size_t k = 0; k < mul.inputs.size() && !sc; ++k) sc = m.init_of(mul, k);
      if (!sc || sc->data.size() != 1) { m.fail("Mul after " + h + ".reg is not a scalar Scale"); return bad(); }
      for (float& v : w.tensors[w.index.at(h + ".reg.w")].data) v *= sc->data[0];
      for (float& v : w.tensors[w.index.at(h + ".reg.b")].data) v *= sc->data[0];
    }
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
