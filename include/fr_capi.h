/*
 * fr_capi.h -- the C ABI of the B200-native face pipeline.
 *
 * This is the drop-in boundary for the hot path of cucibala/FaceRecognizeOnnx:
 * everything the reference does between `FaceDetector::detect` /
 * `FaceRecognizer::extractFeature|compareFaces` and the two
 * `Ort::Session::Run` calls.  POD only: plain pointers, sizes and ints; no
 * C++/torch/OpenCV types.  `include/face_detector.h` and
 * `include/face_recognizer.h` re-create the reference's C++ classes on top of
 * these entry points; `INTEGRATION.md` shows the binding a maintainer adds.
 *
 * Conventions
 *   - every function returns FR_OK (0) or a negative fr_status; the message is
 *     available from fr_last_error(ctx).  Mirrors the reference's "return
 *     false / empty vector / 0.0f and print on cerr" convention
 *     (src/face_detector.cpp:86-89,142-167,217-219).
 *   - images are 8-bit 3-channel BGR, row stride `step` bytes (cv::Mat::step).
 *   - `memspace` tells whether the *payload* pointers (pixels, outputs) are
 *     host (FR_MEM_HOST) or device (FR_MEM_DEVICE) memory.  Pointer/size
 *     arrays themselves are always host arrays.
 *   - a ctx owns one CUDA stream; calls on one ctx serialise.  There is no
 *     CPU fallback anywhere: without a CUDA device fr_create fails.
 */
#ifndef FR_CAPI_H_
#define FR_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_API __attribute__((visibility("default")))

typedef enum fr_status {
  FR_OK = 0,
  FR_ERR_INVALID_ARG = -1,   /* null pointer, empty image, bad dims (face_detector.cpp:148-156) */
  FR_ERR_NOT_LOADED = -2,    /* model not loaded (face_detector.cpp:142-145) */
  FR_ERR_CUDA = -3,          /* CUDA runtime / driver failure (the Ort::Exception analogue) */
  FR_ERR_MODEL = -4,         /* weight file unreadable / shape mismatch (loadModel -> false) */
  FR_ERR_ALIGN = -5,         /* alignment failed (face_recognizer.cpp:255-258) */
  FR_ERR_CAPACITY = -6,      /* caller buffer too small */
  FR_ERR_UNSUPPORTED = -7,
  FR_ERR_IO = -8             /* gallery file unreadable / unwritable / malformed */
} fr_status;

enum { FR_MEM_HOST = 0, FR_MEM_DEVICE = 1 };
enum { FR_MODEL_DET = 0, FR_MODEL_REC = 1 };
enum { FR_DET_SIZE = 640, FR_REC_SIZE = 112, FR_FEAT_DIM = 512, FR_NUM_ANCHORS = 16800 };

/* Layout-identical to the reference's
 *   struct FaceBox { cv::Rect box; float score; cv::Point2f landmarks[5]; }
 * (src/face_detector.h:8-12): 4 x int32, float, 10 x float = 60 bytes. */
typedef struct fr_face {
  int32_t x, y, w, h;
  float score;
  float lm[10];
} fr_face;

typedef struct fr_ctx fr_ctx;
typedef struct fr_weights fr_weights;
typedef struct fr_gallery fr_gallery;

/* ---------------------------------------------------------------- weights --
 * Host-only (no GPU needed).  Replaces what `loadModel` reads from the .onnx
 * file (src/face_detector.cpp:20-90, src/face_recognizer.cpp:21-91).  With
 * onnx_path == NULL (or unreadable -> FR_ERR_MODEL) the tensors are seeded
 * random-init weights of the same architecture.  Tensors are in canonical
 * layout: conv OIHW with BN folded, fc [out,in], affine (scale, shift). */
FR_API int fr_weights_create(fr_weights** out, int model, const char* onnx_path, uint64_t seed);
FR_API void fr_weights_destroy(fr_weights* w);
FR_API int fr_weights_model(const fr_weights* w);
FR_API int fr_weights_from_onnx(const fr_weights* w); /* 1 if loaded from file */
FR_API int fr_weights_num_tensors(const fr_weights* w);
FR_API int fr_weights_tensor_info(const fr_weights* w, int idx, char* name, int name_cap,
                                  int64_t dims[4], int* ndim);
FR_API int fr_weights_tensor_get(const fr_weights* w, int idx, float* out, size_t n);
FR_API int fr_weights_tensor_set(fr_weights* w, int idx, const float* data, size_t n);
FR_API const char* fr_weights_last_error(void);

/* -------------------------------------------------------------- lifecycle -- */
FR_API int fr_create(fr_ctx** out, int device, const fr_weights* det, const fr_weights* rec);
FR_API void fr_destroy(fr_ctx* ctx);
FR_API const char* fr_last_error(const fr_ctx* ctx);
/* Use an external cudaStream_t (e.g. torch's current stream); NULL restores the ctx-owned one. */
FR_API int fr_set_stream(fr_ctx* ctx, void* cuda_stream);
FR_API int fr_synchronize(fr_ctx* ctx);
/* Number of kernels this library launched on ctx since creation (bench "gpu_launches"). */
FR_API uint64_t fr_launch_count(const fr_ctx* ctx);

/* ------------------------------------------------------------ detection ---
 * Replaces FaceDetector::detect (src/face_detector.h:20, .cpp:139-222):
 * letterbox preprocess -> SCRFD -> three-stride decode -> threshold ->
 * /scale -> int rects -> integer-IoU greedy NMS.  Faces come back sorted by
 * score descending (the order nms() leaves them in, .cpp:376-383).
 * out holds cap_per_img records per image; n_out[i] = number written. */
FR_API int fr_detect(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                     float score_thr, float nms_thr, fr_face* out, int cap, int* n_out);
FR_API int fr_detect_batch(fr_ctx* ctx, const uint8_t* const* bgr, const int* rows, const int* cols,
                           const size_t* step, int n_img, int memspace, float score_thr,
                           float nms_thr, fr_face* out, int cap_per_img, int* n_out);

/* ------------------------------------------------------------ recognition --
 * Replaces FaceRecognizer::extractFeature (src/face_recognizer.h:15,
 * .cpp:236-304): estimateAffinePartial2D(landmarks -> template) -> warpAffine
 * 112x112 (or crop+resize fallback, .cpp:116-127) -> BGR2RGB, (v-127.5)/128 ->
 * IResNet-50 -> L2 normalise.  valid[i] = 0 marks the reference's "empty
 * vector" outcome (alignment impossible); its output row is zero-filled. */
FR_API int fr_embed(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                    const fr_face* face, float* out512);
FR_API int fr_embed_faces_batch(fr_ctx* ctx, const uint8_t* const* bgr, const int* rows,
                                const int* cols, const size_t* step, int n_img, int memspace,
                                const fr_face* faces, const int* face_img, int n_faces,
                                float* out /* [n_faces,512] */, int* valid /* [n_faces] */);
/* extractFeatureSimple (src/face_recognizer.cpp:152-234): plain resize to 112x112. */
FR_API int fr_embed_simple(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                           float* out512);
/* Already aligned 112x112x3 BGR crops, contiguous [n,112,112,3]. */
FR_API int fr_embed_aligned_batch(fr_ctx* ctx, const uint8_t* crops, int n, int memspace,
                                  float* out /* [n,512] */);

/* compareFaces (src/face_recognizer.cpp:320-334): (dot+1)/2, 0.0f on size
 * mismatch / empty.  Pure host arithmetic on two host vectors, sequential fp32
 * accumulation exactly as the reference. */
FR_API float fr_compare(const float* a, int dim_a, const float* b, int dim_b);
/* Batched 1:1 on the GPU: out[i] = (dot(a_i, b_i) + 1) / 2. */
FR_API int fr_compare_batch(fr_ctx* ctx, const float* a, const float* b, int n, int dim,
                            int memspace, float* out);

/* ------------------------------------------------------- fused pipeline ----
 * det + align + embed for a batch of frames with a fixed number K of faces
 * per frame (static shapes, no host round trip between stages): slot j of
 * frame i is the j-th post-NMS detection if it exists, otherwise
 * pad_faces[i*K+j] (if pad_faces != NULL) or invalid.  This is the shape the
 * reference's webcam loop generalises to (src/main.cpp:214-258). */
FR_API int fr_pipeline_batch(fr_ctx* ctx, const uint8_t* const* bgr, const int* rows,
                             const int* cols, const size_t* step, int n_img, int memspace,
                             float score_thr, float nms_thr, int faces_per_img,
                             const fr_face* pad_faces, fr_face* out_faces, int* out_n_det,
                             float* out_emb, int* out_valid);
/* Asynchronous form (host buffers only): enqueue the batch and return a ticket.  The frames are
 * copied on a separate copy stream into one of two staging slots, so the upload of batch i+1
 * overlaps the compute of batch i; outputs are valid after fr_pipeline_wait(ticket).  At most
 * two batches may be in flight; input and output buffers must stay alive (and should be
 * page-locked for the copies to be asynchronous) until the wait returns.  The results return on a
 * third stream, so batch i+1 starts computing as soon as the last kernel of batch i has finished.
 *
 * Launch chains whose arguments repeat (this call, fr_pipeline_batch on device-resident frames,
 * fr_detect(_batch) / fr_embed(_faces_batch) on up to 4 images) are replayed as CUDA graphs from
 * the third identical call on; the results are the same bytes either way.  FR_GRAPHS=0 in the
 * environment turns that off. */
FR_API int fr_pipeline_submit(fr_ctx* ctx, const uint8_t* const* bgr, const int* rows,
                              const int* cols, const size_t* step, int n_img, float score_thr,
                              float nms_thr, int faces_per_img, const fr_face* pad_faces,
                              fr_face* out_faces, int* out_n_det, float* out_emb, int* out_valid,
                              int* ticket);
FR_API int fr_pipeline_wait(fr_ctx* ctx, int ticket);

/* ------------------------------------------------------------- 1:N search --
 * North-star extension (the reference has only 1:1, src/face_recognizer.cpp:320-334).
 * Gallery rows are L2-normalised fp32 in, stored bf16 on the device.  search
 * returns, per query, the top-k raw cosine scores (descending; ties -> lower
 * index) and global row indices (index_base + local row). */
FR_API int fr_gallery_create(fr_ctx* ctx, fr_gallery** out, int64_t capacity_rows, int64_t index_base);
/* Storage options.  FR_GALLERY_FP8 keeps an e4m3 mirror of the rows (512 B / row) for
 * fr_gallery_search_fp8: a coarse pass on the fp8 tensor cores (kind::f8f6f4, twice the bf16 rate,
 * half the bytes) followed by an EXACT bf16 re-rank of the union of the per-split top-16 candidate
 * lists (64 candidates per query at 4096 queries, up to 128).  With FR_GALLERY_BF16_ON_HOST the bf16
 * rows that only the re-rank touches live in mapped pinned host memory, so a GPU holds twice the
 * rows per GB of HBM; the plain bf16 search is then unavailable (FR_ERR_UNSUPPORTED). */
enum { FR_GALLERY_BF16 = 0, FR_GALLERY_FP8 = 1, FR_GALLERY_BF16_ON_HOST = 2 };
FR_API int fr_gallery_create_ex(fr_ctx* ctx, fr_gallery** out, int64_t capacity_rows,
                                int64_t index_base, int flags);
FR_API void fr_gallery_destroy(fr_gallery* g);
FR_API int fr_gallery_add(fr_gallery* g, const float* rows, int64_t n, int memspace);
/* Fill with n synthetic unit-norm rows generated on the device (Philox-style hash of seed,row). */
FR_API int fr_gallery_fill_synthetic(fr_gallery* g, int64_t n, uint64_t seed);
FR_API int fr_gallery_get_rows(fr_gallery* g, int64_t first, int64_t n, float* out_host);
FR_API int64_t fr_gallery_size(const fr_gallery* g);
/* Persistence / enrolment: save a shard (bf16 rows as stored, bit-exact round trip), append the
 * rows of a saved shard, remove one row (the last row moves into its place). */
FR_API int fr_gallery_save(fr_gallery* g, const char* path);
FR_API int fr_gallery_load(fr_gallery* g, const char* path, int64_t* file_index_base);
FR_API int fr_gallery_remove(fr_gallery* g, int64_t row);
FR_API int fr_gallery_search(fr_gallery* g, const float* queries, int nq, int k, int memspace,
                             float* out_scores, int64_t* out_idx);
/* fp8 coarse pass + exact bf16 re-rank (needs FR_GALLERY_FP8).  Every returned score is the bf16
 * search's score of the returned row; the returned set equals the bf16 search's wherever the true
 * top-k members rank inside their split's top 16 by fp8 score (fp8 score error ~ 4e-3 rms). */
FR_API int fr_gallery_search_fp8(fr_gallery* g, const float* queries, int nq, int k, int memspace,
                                 float* out_scores, int64_t* out_idx);
/* Merge `parts` per-shard top-k lists [parts][nq][k] (as gathered over NCCL) into [nq][k]. */
FR_API int fr_topk_merge(fr_ctx* ctx, const float* scores, const int64_t* idx, int parts, int nq,
                         int k, int memspace, float* out_scores, int64_t* out_idx);
/* Row-sharded search across GPUs (one rank per GPU, each holding rows [index_base, index_base +
 * size) of the global gallery).  Every rank calls this with the same queries:
 * local fused GEMM + top-k with global indices -> ONE ncclAllGather of packed 8-byte
 * {fp32 score, uint32 global index} records over NVLink -> rank merge (score desc, global index
 * asc) on every rank, so the result is identical on all ranks and independent of `world`.
 * `nccl_comm` is the caller's ncclComm_t for ctx's device (void* so this header needs no nccl.h);
 * the library resolves ncclAllGather from the NCCL already in the process (else libnccl.so.2).
 * Asynchronous on ctx's stream when memspace == FR_MEM_DEVICE. */
FR_API int fr_gallery_search_sharded(fr_gallery* g, void* nccl_comm, int world, const float* queries,
                                     int nq, int k, int memspace, float* out_scores,
                                     int64_t* out_idx);
/* The two halves of the above for hosts whose collective layer is not raw NCCL (e.g.
 * torch.distributed): the local search emitting packed records [nq][k], and the merge of
 * gathered records [parts][nq][k]. */
FR_API int fr_gallery_search_packed(fr_gallery* g, const float* queries, int nq, int k, int memspace,
                                    uint64_t* out_records);
FR_API int fr_gallery_search_packed_fp8(fr_gallery* g, const float* queries, int nq, int k,
                                        int memspace, uint64_t* out_records);
FR_API int fr_topk_merge_packed(fr_ctx* ctx, const uint64_t* records, int parts, int nq, int k,
                                int memspace, float* out_scores, int64_t* out_idx);

/* -------------------------------------------------------- instrumentation --
 * CUDA-event timing of each stage on the ctx stream (what bench.py's roofline block uses).
 * fr_stage_times synchronises, adds the elapsed ms of every recorded interval to ms[stage]
 * (accumulating since the last reset) and clears the interval list. */
enum {
  FR_STAGE_PREPROCESS = 0, /* K1 */
  FR_STAGE_SCRFD = 1,      /* K2 */
  FR_STAGE_DECODE_NMS = 2, /* K3+K4 */
  FR_STAGE_ALIGN = 3,      /* K5 */
  FR_STAGE_STEM = 4,       /* K6 stem (SIMT) */
  FR_STAGE_TRUNK = 5,      /* K6/K7 tcgen05 shift-GEMM launches */
  FR_STAGE_L2NORM = 6,     /* R4 */
  FR_STAGE_GALLERY = 7,    /* K9 */
  FR_NUM_STAGES = 8
};
FR_API int fr_enable_stage_timing(fr_ctx* ctx, int on);
FR_API int fr_stage_times(fr_ctx* ctx, double ms[FR_NUM_STAGES], int reset);

/* ------------------------------------------------------ stage-level hooks --
 * Each stage of the hot path, callable alone, for the parity tests. */
/* K1: FaceDetector::preprocess (src/face_detector.cpp:92-137). out_chw fp32 [n,3,640,640]. */
FR_API int fr_det_preprocess(fr_ctx* ctx, const uint8_t* const* bgr, const int* rows,
                             const int* cols, const size_t* step, int n_img, int memspace,
                             float* out_chw, float* out_scale);
/* K2: SCRFD forward on a prepared fp32 CHW batch (host). heads[9]: scores x3 [n,N_s],
 * bbox x3 [n,N_s,4], kps x3 [n,N_s,10] (host). */
FR_API int fr_scrfd_forward(fr_ctx* ctx, const float* chw, int n_img, float* const heads[9]);
/* K3+K4: decode + postprocess + NMS on caller-supplied head tensors (host, layout as above). */
FR_API int fr_scrfd_decode_nms(fr_ctx* ctx, const float* const heads[9], int n_img,
                               const float* scales, float score_thr, float nms_thr,
                               fr_face* out, int cap_per_img, int* n_out);
/* K5a: similarity estimate (RANSAC + least squares). M_out [n,6] double, ok [n]. */
FR_API int fr_estimate_alignment(fr_ctx* ctx, const float* landmarks /* [n,10] */, int n,
                                 double* M_out, int* ok);
/* K5b: alignFace for faces of ONE image -> [n,112,112,3] BGR u8 crops (host). */
FR_API int fr_align_faces(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                          const fr_face* faces, int n_faces, uint8_t* out_crops, int* valid);
/* K5c: cv::warpAffine with a caller-supplied forward matrix (double[6]). */
FR_API int fr_warp_affine(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                          const double* M, uint8_t* out_crop);
/* cv::resize(INTER_LINEAR) of a u8 BGR image to (new_w,new_h), host in/out. */
FR_API int fr_resize_linear(fr_ctx* ctx, const uint8_t* bgr, int rows, int cols, size_t step,
                            int new_w, int new_h, uint8_t* out);
/* K6/K7: IResNet-50 on fp32 CHW RGB [-1,1] input (host) -> raw [n,512] (not normalised). */
FR_API int fr_iresnet_forward(fr_ctx* ctx, const float* chw, int n, float* out_raw);
/* Activation tap after block `tap` of the last IResNet forward (0 = stem, 1..24 = blocks),
 * returned as fp32 NCHW (host). */
FR_API int fr_iresnet_tap(fr_ctx* ctx, int tap, int n, float* out, size_t out_elems);
/* Activation tap of the last SCRFD forward, returned as fp32 NHWC (host).  tap: 0 stem, 1 b0,
 * 2..14 backbone blocks, 15..17 laterals, 18..20 fpn/inter, 21..22 pafpn outs of strides 16/32,
 * 23..25 / 26..28 head towers. */
FR_API int fr_scrfd_tap(fr_ctx* ctx, int tap, int n, float* out, size_t out_elems);
/* R4: L2 normalise rows (src/face_recognizer.cpp:306-318). */
FR_API int fr_l2_normalize(fr_ctx* ctx, const float* in, int n, int dim, int memspace, float* out);
/* Unit-test hook for the tcgen05 implicit-GEMM conv: x NCHW fp32, w OIHW fp32 (3x3 pad 1 or
 * 1x1), optional pre-affine (scale,shift per input channel), bias, PReLU slope, residual
 * (NCHW fp32, output shape).  stride 1 or 2.  y NCHW fp32 (host). */
FR_API int fr_test_conv(fr_ctx* ctx, const float* x, int n, int cin, int h, int w,
                        const float* wgt, int cout, int ksize, int stride,
                        const float* pre_scale, const float* pre_shift, const float* bias,
                        const float* prelu, const float* residual, float* y);

#ifdef __cplusplus
}
#endif
#endif /* FR_CAPI_H_ */
