// K9: 1:N cosine search (north_star extension; the reference only has 1:1 compareFaces,
// src/face_recognizer.cpp:320-334).
#include "common.h"

struct fr_gallery {
  fr_ctx* ctx = nullptr;
};

extern "C" {
int fr_gallery_create(fr_ctx* ctx, fr_gallery** out, int64_t capacity_rows, int64_t index_base) {
  (void)out; (void)capacity_rows; (void)index_base;
  return fr_fail(ctx, FR_ERR_UNSUPPORTED, "gallery not built yet");
}
void fr_gallery_destroy(fr_gallery* g) { delete g; }
int fr_gallery_add(fr_gallery* g, const float* rows, int64_t n, int memspace) { (void)g; (void)rows; (void)n; (void)memspace; return FR_ERR_UNSUPPORTED; }
int fr_gallery_fill_synthetic(fr_gallery* g, int64_t n, uint64_t seed) { (void)g; (void)n; (void)seed; return FR_ERR_UNSUPPORTED; }
int fr_gallery_get_rows(fr_gallery* g, int64_t first, int64_t n, float* out_host) { (void)g; (void)first; (void)n; (void)out_host; return FR_ERR_UNSUPPORTED; }
int64_t fr_gallery_size(const fr_gallery* g) { (void)g; return 0; }
int fr_gallery_search(fr_gallery* g, const float* queries, int nq, int k, int memspace, float* out_scores, int64_t* out_idx) {
  (void)g; (void)queries; (void)nq; (void)k; (void)memspace; (void)out_scores; (void)out_idx; return FR_ERR_UNSUPPORTED;
}
int fr_topk_merge(fr_ctx* ctx, const float* scores, const int64_t* idx, int parts, int nq, int k, int memspace, float* out_scores, int64_t* out_idx) {
  (void)scores; (void)idx; (void)parts; (void)nq; (void)k; (void)memspace; (void)out_scores; (void)out_idx;
  return fr_fail(ctx, FR_ERR_UNSUPPORTED, "gallery not built yet");
}
}
