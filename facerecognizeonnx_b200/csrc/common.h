// Internal declarations shared by the host runtime and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <set>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fr_capi.h"

// ----------------------------------------------------------------- weights --
struct fr_tensor {
  std::string name;
  std::vector<int64_t> dims;
  std::vector<float> data;
  size_t numel() const {
    size_t n = 1;
    for (auto d : dims) n *= (size_t)d;
    return n;
  }
};

struct fr_weights {
  int model = 0;
  bool from_onnx = false;
  uint64_t seed = 0;
  std::vector<fr_tensor> tensors;
  std::map<std::string, int> index;
  const fr_tensor& at(const std::string& n) const;
  bool has(const std::string& n) const { return index.count(n) != 0; }
};

void fr_weights_build_spec(fr_weights& w);                 // names + dims, zero data
void fr_weights_random_init(fr_weights& w, uint64_t seed); // seeded init of every tensor
int fr_weights_load_onnx(fr_weights& w, const char* path, std::string& err);

// Bumped whenever the library frees or reallocates device memory that kernels of an earlier call may have
// been pointed at: a cached CUDA graph (capi.cu) is only replayed while the epoch it was captured under
// is still the current one.
inline std::atomic<uint64_t>& fr_alloc_epoch() {
  static std::atomic<uint64_t> e{0};
  return e;
}

// --------------------------------------------------------------------- ctx --
struct DetModel;  // k_scrfd.cu
struct RecModel;  // k_iresnet.cu

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grows (never shrinks); contents are not preserved; returns false on OOM
  bool reserve(size_t bytes, bool zero = false);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Image descriptor consumed by the preprocess / warp kernels (device copy).
struct ImgDesc {
  const uint8_t* ptr;
  int rows, cols;
  long long step;
  float scale;     // letterbox scale (fp32, face_detector.cpp:101-103)
  int new_w, new_h;
};

struct NmsScratch {
  DevBuf keys;    // uint64 [n_img][32768]
  DevBuf counts;  // int [n_img]
  DevBuf boxes;   // int4 [n_img][FR_NUM_ANCHORS] (spill path for > 4096 candidates)
};

struct fr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  std::mutex mu;
  uint64_t launches = 0;
  // per-device launch state: cudaFuncSetAttribute and the SM count belong to a device, and a
  // process may hold contexts on several (never cache them in function-level statics)
  int num_sms = 0;
  std::set<const void*> smem_opt_in;
  DetModel* det = nullptr;
  RecModel* rec = nullptr;
  // generic staging buffers (device) + pinned host staging
  DevBuf img_stage;     // uploaded frames
  DevBuf img_desc;      // ImgDesc[n]
  DevBuf faces_dev;     // fr_face scratch
  DevBuf misc[12];
  NmsScratch nms;
  // device-side cache of descriptor sets: a steady-state loop over a few fixed batches never
  // re-uploads (and so never synchronises the host)
  struct DescSlot { std::vector<ImgDesc> h; DevBuf d; uint64_t stamp = 0; };
  std::vector<DescSlot> desc_cache;
  uint64_t desc_stamp = 0;
  // optional per-stage CUDA-event timing (fr_enable_stage_timing / fr_stage_times)
  bool timing = false;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timing_events;
  double stage_ms[FR_NUM_STAGES] = {0};
  void stage_begin(int stage);
  void stage_end();
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  void* pin(size_t bytes);
  // asynchronous pipeline (fr_pipeline_submit / fr_pipeline_wait): the frames of batch i+1 are
  // copied on copy_stream into the other staging slot while batch i computes on `stream`
  struct PipeSlot {
    DevBuf stage;                                  // frames
    DevBuf pad, r_faces, r_ndet, r_emb, r_valid;   // padding faces (H2D) and the batch's results (D2H)
    cudaEvent_t h2d = nullptr, computed = nullptr, done = nullptr;
    bool busy = false;
  };
  cudaStream_t copy_stream = nullptr;              // H2D
  cudaStream_t d2h_stream = nullptr;               // results go back while the next batch already computes
  PipeSlot pslots[2];
  int pslot_next = 0;
  // small-batch launch chains (fr_detect / fr_embed at a handful of images) replayed as CUDA graphs
  struct GraphSlot {
    std::vector<long long> key;      // everything the captured kernel arguments depend on
    cudaGraphExec_t exec = nullptr;
    uint64_t epoch = 0;              // fr_alloc_epoch() the key was last run / captured under
    uint64_t launches = 0;           // kernels per replay (for fr_launch_count)
    int state = 0;                   // 0 new, 1 ran eagerly once, 2 captured, -1 capture failed: stay eager
  };
  std::vector<GraphSlot> graphs;
};

#define FR_CUDA_OK(ctx, expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return FR_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define FR_CHECK(expr)                  \
  do {                                  \
    int _s = (expr);                    \
    if (_s != FR_OK) return _s;         \
  } while (0)

static inline int fr_fail(fr_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to `bytes` of dynamic shared memory on ctx's device, once per ctx.
template <typename F>
static inline cudaError_t fr_opt_in_smem(fr_ctx* ctx, F* func, int bytes) {
  const void* key = reinterpret_cast<const void*>(func);
  if (ctx->smem_opt_in.count(key)) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) ctx->smem_opt_in.insert(key);
  return e;
}

// ------------------------------------------------------------ stage launchers
// k_preprocess.cu
int k_det_preprocess(fr_ctx* ctx, const ImgDesc* d_desc, int n_img, __nv_bfloat16* d_out_chw);
int k_bf16_to_f32(fr_ctx* ctx, const __nv_bfloat16* in, float* out, size_t n);
int k_f32_to_bf16(fr_ctx* ctx, const float* in, __nv_bfloat16* out, size_t n);
int k_resize_u8(fr_ctx* ctx, const uint8_t* src, int rows, int cols, long long step, int new_w,
                int new_h, uint8_t* dst);

// k_decode_nms.cu
struct HeadPtrs {
  const float* score[3];
  const float* bbox[3];
  const float* kps[3];
};
int k_scrfd_decode_nms(fr_ctx* ctx, NmsScratch& s, const HeadPtrs& heads, int n_img,
                       const ImgDesc* d_desc /* scale */, const float* d_scales /* or null */,
                       float score_thr, float nms_thr, fr_face* d_out, int cap_per_img,
                       int* d_n_out);

// k_warp.cu
struct AlignRec {       // per-face alignment record (device)
  double inv[6];        // inverse affine [[A11,A12,b1],[A21,A22,b2]]
  double fwd[6];        // forward matrix M (for the test hook)
  int mode;             // 0 warpAffine, 1 crop+resize fallback, 2 invalid
  int img;              // index into ImgDesc
  int cx, cy, cw, ch;   // crop rect for mode 1
  int pad;
};
int k_align_estimate(fr_ctx* ctx, const fr_face* d_faces, const int* d_face_img, int n_faces,
                     const ImgDesc* d_desc, int n_img, AlignRec* d_rec);
int k_align_select(fr_ctx* ctx, const fr_face* d_det, const int* d_n_det, int cap_per_img,
                   const fr_face* d_pad, int n_img, int k, fr_face* d_sel, int* d_face_img,
                   int* d_valid);
int k_align_warp(fr_ctx* ctx, const AlignRec* d_rec, int n_faces, const ImgDesc* d_desc,
                 uint8_t* d_crops, int* d_valid);

// k_scrfd.cu
int det_model_create(fr_ctx* ctx, const fr_weights* w);
void det_model_destroy(fr_ctx* ctx);
int det_forward(fr_ctx* ctx, const __nv_bfloat16* d_in_chw, int n_img, HeadPtrs* heads);
int det_tap(fr_ctx* ctx, int tap, int n, float* h_out, size_t out_elems);

// k_iresnet.cu
int rec_model_create(fr_ctx* ctx, const fr_weights* w);
void rec_model_destroy(fr_ctx* ctx);
// crops: u8 BGR [n,112,112,3] (device) -> raw fp32 [n,512] (device, d_out_raw) and
// L2-normalised fp32 (d_out_norm, may be null).
int rec_forward_crops(fr_ctx* ctx, const uint8_t* d_crops, int n, float* d_out_raw,
                      float* d_out_norm, const int* d_valid);
// fp32 CHW RGB [-1,1] input (device) for the test hook.
int rec_forward_chw(fr_ctx* ctx, const float* d_chw, int n, float* d_out_raw);
int rec_tap(fr_ctx* ctx, int tap, int n, float* h_out, size_t out_elems);
int rec_test_conv(fr_ctx* ctx, const float* x, int n, int cin, int h, int w, const float* wgt,
                  int cout, int ksize, int stride, const float* pre_scale,
                  const float* pre_shift, const float* bias, const float* prelu,
                  const float* residual, float* y);
int k_l2_normalize(fr_ctx* ctx, const float* d_in, int n, int dim, float* d_out,
                   const int* d_valid);
int k_compare_batch(fr_ctx* ctx, const float* d_a, const float* d_b, int n, int dim, float* d_out);

// k_gallery.cu
int gallery_create(fr_ctx* ctx, fr_gallery** out, int64_t cap, int64_t base);
