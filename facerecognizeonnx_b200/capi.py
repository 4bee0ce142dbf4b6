"""ctypes binding of include/fr_capi.h (libfr_b200.so).

The shared library is the product; this module only marshals numpy arrays /
raw device pointers across the C ABI.  It fails loudly when the library is
missing or was not built -- there is no Python/CPU fallback for any stage.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfr_b200.so")

FR_OK = 0
FR_ERR_INVALID_ARG, FR_ERR_NOT_LOADED, FR_ERR_CUDA, FR_ERR_MODEL = -1, -2, -3, -4
FR_ERR_ALIGN, FR_ERR_CAPACITY, FR_ERR_UNSUPPORTED, FR_ERR_IO = -5, -6, -7, -8
FR_MEM_HOST, FR_MEM_DEVICE = 0, 1
FR_MODEL_DET, FR_MODEL_REC = 0, 1
DET_SIZE, REC_SIZE, FEAT_DIM, NUM_ANCHORS = 640, 112, 512, 16800
HEAD_N = (12800, 3200, 800)
HEAD_C = (1, 4, 10)

FACE_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"),
                       ("score", "<f4"), ("lm", "<f4", (10,))])
assert FACE_DTYPE.itemsize == 60

# every symbol include/fr_capi.h declares (tests/test_capi_host.py checks the .so exports them)
SYMBOLS = [
    "fr_weights_create", "fr_weights_destroy", "fr_weights_model", "fr_weights_from_onnx",
    "fr_weights_num_tensors", "fr_weights_tensor_info", "fr_weights_tensor_get",
    "fr_weights_tensor_set", "fr_weights_last_error",
    "fr_create", "fr_destroy", "fr_last_error", "fr_set_stream", "fr_synchronize", "fr_launch_count",
    "fr_enable_stage_timing", "fr_stage_times",
    "fr_detect", "fr_detect_batch",
    "fr_embed", "fr_embed_faces_batch", "fr_embed_simple", "fr_embed_aligned_batch",
    "fr_compare", "fr_compare_batch", "fr_pipeline_batch", "fr_pipeline_submit", "fr_pipeline_wait",
    "fr_gallery_create", "fr_gallery_destroy", "fr_gallery_add", "fr_gallery_fill_synthetic",
    "fr_gallery_get_rows", "fr_gallery_size", "fr_gallery_search", "fr_topk_merge",
    "fr_gallery_search_sharded", "fr_gallery_search_packed", "fr_topk_merge_packed",
    "fr_gallery_create_ex", "fr_gallery_search_fp8", "fr_gallery_search_packed_fp8",
    "fr_gallery_save", "fr_gallery_load", "fr_gallery_remove",
    "fr_det_preprocess", "fr_scrfd_forward", "fr_scrfd_decode_nms", "fr_estimate_alignment",
    "fr_align_faces", "fr_warp_affine", "fr_resize_linear", "fr_iresnet_forward", "fr_iresnet_tap",
    "fr_l2_normalize", "fr_test_conv", "fr_scrfd_tap",
]


class FrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fr error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no fallback implementation.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
    L.fr_weights_create.argtypes = [C.POINTER(vp), i32, C.c_char_p, u64]
    L.fr_weights_destroy.argtypes = [vp]
    L.fr_weights_destroy.restype = None
    L.fr_weights_num_tensors.argtypes = [vp]
    L.fr_weights_model.argtypes = [vp]
    L.fr_weights_from_onnx.argtypes = [vp]
    L.fr_weights_tensor_info.argtypes = [vp, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i32)]
    L.fr_weights_tensor_get.argtypes = [vp, i32, vp, sz]
    L.fr_weights_tensor_set.argtypes = [vp, i32, vp, sz]
    L.fr_weights_last_error.restype = C.c_char_p
    L.fr_create.argtypes = [C.POINTER(vp), i32, vp, vp]
    L.fr_destroy.argtypes = [vp]
    L.fr_destroy.restype = None
    L.fr_last_error.argtypes = [vp]
    L.fr_last_error.restype = C.c_char_p
    L.fr_set_stream.argtypes = [vp, vp]
    L.fr_synchronize.argtypes = [vp]
    L.fr_launch_count.argtypes = [vp]
    L.fr_launch_count.restype = u64
    L.fr_enable_stage_timing.argtypes = [vp, i32]
    L.fr_stage_times.argtypes = [vp, vp, i32]
    L.fr_detect.argtypes = [vp, vp, i32, i32, sz, f32, f32, vp, i32, C.POINTER(i32)]
    L.fr_detect_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, f32, vp, i32, vp]
    L.fr_embed.argtypes = [vp, vp, i32, i32, sz, vp, vp]
    L.fr_embed_faces_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, vp]
    L.fr_embed_simple.argtypes = [vp, vp, i32, i32, sz, vp]
    L.fr_embed_aligned_batch.argtypes = [vp, vp, i32, i32, vp]
    L.fr_compare.argtypes = [vp, i32, vp, i32]
    L.fr_compare.restype = f32
    L.fr_compare_batch.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.fr_pipeline_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, f32, i32, vp, vp, vp, vp, vp]
    L.fr_pipeline_submit.argtypes = [vp, vp, vp, vp, vp, i32, f32, f32, i32, vp, vp, vp, vp, vp, vp]
    L.fr_pipeline_wait.argtypes = [vp, i32]
    L.fr_gallery_create.argtypes = [vp, C.POINTER(vp), i64, i64]
    L.fr_gallery_destroy.argtypes = [vp]
    L.fr_gallery_destroy.restype = None
    L.fr_gallery_add.argtypes = [vp, vp, i64, i32]
    L.fr_gallery_fill_synthetic.argtypes = [vp, i64, u64]
    L.fr_gallery_get_rows.argtypes = [vp, i64, i64, vp]
    L.fr_gallery_size.argtypes = [vp]
    L.fr_gallery_size.restype = i64
    L.fr_gallery_save.argtypes = [vp, C.c_char_p]
    L.fr_gallery_load.argtypes = [vp, C.c_char_p, vp]
    L.fr_gallery_remove.argtypes = [vp, i64]
    L.fr_gallery_search.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.fr_topk_merge.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    L.fr_gallery_search_sharded.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp]
    L.fr_gallery_search_packed.argtypes = [vp, vp, i32, i32, i32, vp]
    L.fr_topk_merge_packed.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.fr_gallery_create_ex.argtypes = [vp, C.POINTER(vp), i64, i64, i32]
    L.fr_gallery_search_fp8.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.fr_gallery_search_packed_fp8.argtypes = [vp, vp, i32, i32, i32, vp]
    L.fr_det_preprocess.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.fr_scrfd_forward.argtypes = [vp, vp, i32, vp]
    L.fr_scrfd_decode_nms.argtypes = [vp, vp, i32, vp, f32, f32, vp, i32, vp]
    L.fr_estimate_alignment.argtypes = [vp, vp, i32, vp, vp]
    L.fr_align_faces.argtypes = [vp, vp, i32, i32, sz, vp, i32, vp, vp]
    L.fr_warp_affine.argtypes = [vp, vp, i32, i32, sz, vp, vp]
    L.fr_resize_linear.argtypes = [vp, vp, i32, i32, sz, i32, i32, vp]
    L.fr_iresnet_forward.argtypes = [vp, vp, i32, vp]
    L.fr_iresnet_tap.argtypes = [vp, i32, i32, vp, sz]
    L.fr_scrfd_tap.argtypes = [vp, i32, i32, vp, sz]
    L.fr_l2_normalize.argtypes = [vp, vp, i32, i32, i32, vp]
    L.fr_test_conv.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


# ---------------------------------------------------------------- weights --
class Weights:
    """Host-only canonical weight store (fr_weights_*)."""

    def __init__(self, model: int, onnx_path: Optional[str] = None, seed: int = 0):
        L = lib()
        h = C.c_void_p()
        s = L.fr_weights_create(C.byref(h), model, onnx_path.encode() if onnx_path else None, seed)
        if s != FR_OK:
            raise FrError(s, L.fr_weights_last_error().decode())
        self.h = h
        self.model = model

    def close(self):
        if getattr(self, "h", None):
            lib().fr_weights_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def from_onnx(self) -> bool:
        return bool(lib().fr_weights_from_onnx(self.h))

    def specs(self):
        L = lib()
        out = []
        for i in range(L.fr_weights_num_tensors(self.h)):
            name = C.create_string_buffer(128)
            dims = (C.c_int64 * 4)()
            nd = C.c_int()
            L.fr_weights_tensor_info(self.h, i, name, 128, dims, C.byref(nd))
            out.append((name.value.decode(), tuple(int(dims[k]) for k in range(nd.value))))
        return out

    def to_dict(self):
        L = lib()
        d = {}
        for i, (name, shape) in enumerate(self.specs()):
            a = np.empty(shape, np.float32)
            s = L.fr_weights_tensor_get(self.h, i, a.ctypes.data, a.size)
            if s != FR_OK:
                raise FrError(s, "tensor_get " + name)
            d[name] = a
        return d

    def set(self, name: str, value: np.ndarray):
        for i, (n, shape) in enumerate(self.specs()):
            if n == name:
                a = np.ascontiguousarray(value, np.float32).reshape(shape)
                s = lib().fr_weights_tensor_set(self.h, i, a.ctypes.data, a.size)
                if s != FR_OK:
                    raise FrError(s, "tensor_set " + name)
                return
        raise KeyError(name)


# -------------------------------------------------------------------- ctx --
class _ImageBatch:
    """Marshals a list of HxWx3 uint8 BGR images (numpy, possibly strided rows)
    or raw device pointers into the (ptr[], rows[], cols[], step[]) arrays."""

    def __init__(self, images: Sequence, memspace: int, shapes=None):
        n = len(images)
        self.n = n
        self.ptrs = (C.c_void_p * n)()
        self.rows = (C.c_int * n)()
        self.cols = (C.c_int * n)()
        self.step = (C.c_size_t * n)()
        self.keep = []
        for i, im in enumerate(images):
            if memspace == FR_MEM_DEVICE:
                ptr, r, c, st = im if isinstance(im, tuple) else (im, *shapes[i])
                self.ptrs[i], self.rows[i], self.cols[i], self.step[i] = ptr, r, c, st
            else:
                assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3
                if im.strides[2] != 1 or im.strides[1] != 3:
                    im = np.ascontiguousarray(im)
                self.keep.append(im)
                self.ptrs[i] = im.ctypes.data
                self.rows[i], self.cols[i] = im.shape[0], im.shape[1]
                self.step[i] = im.strides[0]


class Context:
    def __init__(self, device: int = 0, det: Optional[Weights] = None, rec: Optional[Weights] = None):
        L = lib()
        h = C.c_void_p()
        s = L.fr_create(C.byref(h), device, det.h if det else None, rec.h if rec else None)
        if s != FR_OK:
            raise FrError(s, "fr_create failed (needs a CUDA sm_100 device; there is no CPU path)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().fr_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, s: int):
        if s != FR_OK:
            raise FrError(s, lib().fr_last_error(self.h).decode())

    def set_stream(self, stream_ptr: Optional[int]):
        self._check(lib().fr_set_stream(self.h, stream_ptr))

    def synchronize(self):
        self._check(lib().fr_synchronize(self.h))

    def launch_count(self) -> int:
        return int(lib().fr_launch_count(self.h))

    STAGES = ("preprocess", "scrfd", "decode_nms", "align", "stem", "trunk", "l2norm", "gallery")

    def enable_stage_timing(self, on: bool = True):
        self._check(lib().fr_enable_stage_timing(self.h, int(on)))

    def stage_times(self, reset: bool = True) -> dict:
        ms = (C.c_double * 8)()
        self._check(lib().fr_stage_times(self.h, ms, int(reset)))
        return {k: float(ms[i]) for i, k in enumerate(self.STAGES)}

    # -- detection
    def detect_batch(self, images: Sequence[np.ndarray], score_thr=0.5, nms_thr=0.4, cap=256):
        b = _ImageBatch(images, FR_MEM_HOST)
        out = np.zeros((b.n, cap), FACE_DTYPE)
        n_out = np.zeros(b.n, np.int32)
        self._check(lib().fr_detect_batch(self.h, b.ptrs, b.rows, b.cols, b.step, b.n, FR_MEM_HOST,
                                          score_thr, nms_thr, out.ctypes.data, cap, n_out.ctypes.data))
        return [out[i, :n_out[i]].copy() for i in range(b.n)]

    def detect(self, image: np.ndarray, score_thr=0.5, nms_thr=0.4, cap=256):
        return self.detect_batch([image], score_thr, nms_thr, cap)[0]

    # -- recognition
    def embed_faces(self, images: Sequence[np.ndarray], faces: np.ndarray, face_img: Sequence[int]):
        b = _ImageBatch(images, FR_MEM_HOST)
        faces = np.ascontiguousarray(faces, FACE_DTYPE)
        fi = np.ascontiguousarray(face_img, np.int32)
        n = faces.shape[0]
        out = np.zeros((n, FEAT_DIM), np.float32)
        valid = np.zeros(n, np.int32)
        self._check(lib().fr_embed_faces_batch(self.h, b.ptrs, b.rows, b.cols, b.step, b.n, FR_MEM_HOST,
                                               faces.ctypes.data, fi.ctypes.data, n, out.ctypes.data,
                                               valid.ctypes.data))
        return out, valid

    def embed_simple(self, image: np.ndarray) -> np.ndarray:
        b = _ImageBatch([image], FR_MEM_HOST)
        out = np.zeros(FEAT_DIM, np.float32)
        self._check(lib().fr_embed_simple(self.h, b.ptrs[0], b.rows[0], b.cols[0], b.step[0], out.ctypes.data))
        return out

    def embed_aligned(self, crops: np.ndarray) -> np.ndarray:
        crops = np.ascontiguousarray(crops, np.uint8)
        n = crops.shape[0]
        assert crops.shape[1:] == (REC_SIZE, REC_SIZE, 3)
        out = np.zeros((n, FEAT_DIM), np.float32)
        self._check(lib().fr_embed_aligned_batch(self.h, crops.ctypes.data, n, FR_MEM_HOST, out.ctypes.data))
        return out

    def embed_aligned_dev(self, crops_ptr: int, n: int, out_ptr: int):
        self._check(lib().fr_embed_aligned_batch(self.h, crops_ptr, n, FR_MEM_DEVICE, out_ptr))

    def compare_batch(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros(a.shape[0], np.float32)
        self._check(lib().fr_compare_batch(self.h, a.ctypes.data, b.ctypes.data, a.shape[0], a.shape[1],
                                           FR_MEM_HOST, out.ctypes.data))
        return out

    # -- fused pipeline
    def pipeline(self, images: Sequence[np.ndarray], faces_per_img: int, pad_faces: Optional[np.ndarray] = None,
                 score_thr=0.5, nms_thr=0.4):
        b = _ImageBatch(images, FR_MEM_HOST)
        K = faces_per_img
        faces = np.zeros((b.n, K), FACE_DTYPE)
        n_det = np.zeros(b.n, np.int32)
        emb = np.zeros((b.n, K, FEAT_DIM), np.float32)
        valid = np.zeros((b.n, K), np.int32)
        pad = np.ascontiguousarray(pad_faces, FACE_DTYPE) if pad_faces is not None else None
        self._check(lib().fr_pipeline_batch(self.h, b.ptrs, b.rows, b.cols, b.step, b.n, FR_MEM_HOST,
                                            score_thr, nms_thr, K, _ptr(pad), faces.ctypes.data,
                                            n_det.ctypes.data, emb.ctypes.data, valid.ctypes.data))
        return faces, n_det, emb, valid

    def pipeline_submit(self, images: Sequence[np.ndarray], faces_per_img: int, pad_faces: Optional[np.ndarray],
                        out_faces: np.ndarray, out_n_det: np.ndarray, out_emb: np.ndarray, out_valid: np.ndarray,
                        score_thr=0.5, nms_thr=0.4) -> int:
        """Asynchronous pipeline: enqueue a batch (frames + caller-owned, ideally page-locked, output
        arrays) and return a ticket for pipeline_wait.  The upload of this batch overlaps the
        compute of the previous one.  All arrays must stay alive until the wait returns."""
        b = _ImageBatch(images, FR_MEM_HOST)
        t = C.c_int(-1)
        self._inflight = getattr(self, "_inflight", {})
        self._check(lib().fr_pipeline_submit(self.h, b.ptrs, b.rows, b.cols, b.step, b.n, score_thr, nms_thr,
                                             faces_per_img, _ptr(pad_faces), out_faces.ctypes.data,
                                             out_n_det.ctypes.data, out_emb.ctypes.data, out_valid.ctypes.data,
                                             C.byref(t)))
        self._inflight[t.value] = (b, images, pad_faces, out_faces, out_n_det, out_emb, out_valid)
        return t.value

    def pipeline_wait(self, ticket: int) -> None:
        self._check(lib().fr_pipeline_wait(self.h, ticket))
        getattr(self, "_inflight", {}).pop(ticket, None)

    def pipeline_dev(self, frame_ptrs: Sequence[int], rows: int, cols: int, step: int, faces_per_img: int,
                     pad_ptr: Optional[int], out_faces_ptr: Optional[int], out_ndet_ptr: Optional[int],
                     out_emb_ptr: int, out_valid_ptr: Optional[int], score_thr=0.5, nms_thr=0.4):
        n = len(frame_ptrs)
        b = _ImageBatch([(p, rows, cols, step) for p in frame_ptrs], FR_MEM_DEVICE)
        self._check(lib().fr_pipeline_batch(self.h, b.ptrs, b.rows, b.cols, b.step, n, FR_MEM_DEVICE,
                                            score_thr, nms_thr, faces_per_img, pad_ptr, out_faces_ptr,
                                            out_ndet_ptr, out_emb_ptr, out_valid_ptr))

    # -- stage hooks
    def det_preprocess(self, images: Sequence[np.ndarray]):
        b = _ImageBatch(images, FR_MEM_HOST)
        out = np.zeros((b.n, 3, DET_SIZE, DET_SIZE), np.float32)
        scale = np.zeros(b.n, np.float32)
        self._check(lib().fr_det_preprocess(self.h, b.ptrs, b.rows, b.cols, b.step, b.n, FR_MEM_HOST,
                                            out.ctypes.data, scale.ctypes.data))
        return out, scale

    def scrfd_forward(self, chw: np.ndarray) -> List[np.ndarray]:
        chw = np.ascontiguousarray(chw, np.float32)
        n = chw.shape[0]
        heads = [np.zeros((n, HEAD_N[s], HEAD_C[k]), np.float32) for k in range(3) for s in range(3)]
        arr = (C.c_void_p * 9)(*[h.ctypes.data for h in heads])
        self._check(lib().fr_scrfd_forward(self.h, chw.ctypes.data, n, arr))
        return heads

    def scrfd_decode_nms(self, heads: Sequence[np.ndarray], scales, score_thr=0.5, nms_thr=0.4, cap=256):
        hs = [np.ascontiguousarray(h, np.float32) for h in heads]
        n = hs[0].shape[0]
        for k in range(3):
            for s in range(3):
                assert hs[k * 3 + s].size == n * HEAD_N[s] * HEAD_C[k]
        arr = (C.c_void_p * 9)(*[h.ctypes.data for h in hs])
        sc = np.ascontiguousarray(scales, np.float32)
        out = np.zeros((n, cap), FACE_DTYPE)
        n_out = np.zeros(n, np.int32)
        self._check(lib().fr_scrfd_decode_nms(self.h, arr, n, sc.ctypes.data, score_thr, nms_thr,
                                              out.ctypes.data, cap, n_out.ctypes.data))
        return [out[i, :n_out[i]].copy() for i in range(n)]

    def estimate_alignment(self, landmarks: np.ndarray):
        lm = np.ascontiguousarray(landmarks, np.float32).reshape(-1, 10)
        n = lm.shape[0]
        M = np.zeros((n, 2, 3), np.float64)
        ok = np.zeros(n, np.int32)
        self._check(lib().fr_estimate_alignment(self.h, lm.ctypes.data, n, M.ctypes.data, ok.ctypes.data))
        return M, ok

    def align_faces(self, image: np.ndarray, faces: np.ndarray):
        b = _ImageBatch([image], FR_MEM_HOST)
        faces = np.ascontiguousarray(faces, FACE_DTYPE)
        n = faces.shape[0]
        crops = np.zeros((n, REC_SIZE, REC_SIZE, 3), np.uint8)
        valid = np.zeros(n, np.int32)
        self._check(lib().fr_align_faces(self.h, b.ptrs[0], b.rows[0], b.cols[0], b.step[0], faces.ctypes.data,
                                         n, crops.ctypes.data, valid.ctypes.data))
        return crops, valid

    def warp_affine(self, image: np.ndarray, M: np.ndarray) -> np.ndarray:
        b = _ImageBatch([image], FR_MEM_HOST)
        M = np.ascontiguousarray(M, np.float64).reshape(6)
        out = np.zeros((REC_SIZE, REC_SIZE, 3), np.uint8)
        self._check(lib().fr_warp_affine(self.h, b.ptrs[0], b.rows[0], b.cols[0], b.step[0], M.ctypes.data,
                                         out.ctypes.data))
        return out

    def resize_linear(self, image: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
        b = _ImageBatch([image], FR_MEM_HOST)
        out = np.zeros((new_h, new_w, 3), np.uint8)
        self._check(lib().fr_resize_linear(self.h, b.ptrs[0], b.rows[0], b.cols[0], b.step[0], new_w, new_h,
                                           out.ctypes.data))
        return out

    def iresnet_forward(self, chw: np.ndarray) -> np.ndarray:
        chw = np.ascontiguousarray(chw, np.float32)
        n = chw.shape[0]
        out = np.zeros((n, FEAT_DIM), np.float32)
        self._check(lib().fr_iresnet_forward(self.h, chw.ctypes.data, n, out.ctypes.data))
        return out

    def iresnet_tap(self, tap: int, n: int, shape) -> np.ndarray:
        out = np.zeros((n,) + tuple(shape), np.float32)
        self._check(lib().fr_iresnet_tap(self.h, tap, n, out.ctypes.data, out.size))
        return out

    def scrfd_tap(self, tap: int, n: int, shape_chw) -> np.ndarray:
        """Activation `tap` of the last SCRFD forward, returned as NCHW (device layout is NHWC)."""
        c, h, w = shape_chw
        out = np.zeros((n, h, w, c), np.float32)
        self._check(lib().fr_scrfd_tap(self.h, tap, n, out.ctypes.data, out.size))
        return np.ascontiguousarray(out.transpose(0, 3, 1, 2))

    def l2_normalize(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        self._check(lib().fr_l2_normalize(self.h, x.ctypes.data, x.shape[0], x.shape[1], FR_MEM_HOST, out.ctypes.data))
        return out

    def test_conv(self, x, w, stride=1, pre_scale=None, pre_shift=None, bias=None, prelu=None, residual=None):
        x = np.ascontiguousarray(x, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        n, cin, h, wd = x.shape
        cout, _, k, _ = w.shape
        y = np.zeros((n, cout, h // stride, wd // stride), np.float32)
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32)
                for a in (pre_scale, pre_shift, bias, prelu, residual)]
        self._check(lib().fr_test_conv(self.h, x.ctypes.data, n, cin, h, wd, w.ctypes.data, cout, k, stride,
                                       *[_ptr(a) for a in arrs], y.ctypes.data))
        return y


def compare(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)
    return float(lib().fr_compare(a.ctypes.data, a.size, b.ctypes.data, b.size))


# ---------------------------------------------------------------- gallery --
class Gallery:
    """1:N gallery shard on one GPU (fr_gallery_*).  index_base is the global index of local
    row 0, so search results carry global indices."""

    FP8, BF16_ON_HOST = 1, 2     # include/fr_capi.h FR_GALLERY_*

    def __init__(self, ctx: Context, capacity_rows: int, index_base: int = 0, flags: int = 0):
        h = C.c_void_p()
        ctx._check(lib().fr_gallery_create_ex(ctx.h, C.byref(h), capacity_rows, index_base, flags))
        self.h, self.ctx, self.index_base, self.flags = h, ctx, index_base, flags

    def close(self):
        if getattr(self, "h", None):
            lib().fr_gallery_destroy(self.h)
            self.h = None

    __del__ = close

    def __len__(self):
        return int(lib().fr_gallery_size(self.h))

    def add(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, np.float32)
        assert rows.ndim == 2 and rows.shape[1] == FEAT_DIM
        self.ctx._check(lib().fr_gallery_add(self.h, rows.ctypes.data, rows.shape[0], FR_MEM_HOST))

    def fill_synthetic(self, n: int, seed: int):
        self.ctx._check(lib().fr_gallery_fill_synthetic(self.h, n, seed))

    def get_rows(self, first: int, n: int) -> np.ndarray:
        out = np.zeros((n, FEAT_DIM), np.float32)
        self.ctx._check(lib().fr_gallery_get_rows(self.h, first, n, out.ctypes.data))
        return out

    def save(self, path: str):
        self.ctx._check(lib().fr_gallery_save(self.h, path.encode()))

    def load(self, path: str) -> int:
        """Appends the rows of a saved shard; returns the index_base recorded in the file."""
        base = C.c_int64(0)
        self.ctx._check(lib().fr_gallery_load(self.h, path.encode(), C.byref(base)))
        return int(base.value)

    def remove(self, row: int):
        self.ctx._check(lib().fr_gallery_remove(self.h, row))

    def search(self, queries: np.ndarray, k: int = 10):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        s = np.zeros((nq, k), np.float32)
        i = np.zeros((nq, k), np.int64)
        self.ctx._check(lib().fr_gallery_search(self.h, q.ctypes.data, nq, k, FR_MEM_HOST, s.ctypes.data, i.ctypes.data))
        return s, i

    def search_dev(self, q_ptr: int, nq: int, k: int, out_s_ptr: int, out_i_ptr: int):
        self.ctx._check(lib().fr_gallery_search(self.h, q_ptr, nq, k, FR_MEM_DEVICE, out_s_ptr, out_i_ptr))

    def search_fp8(self, queries: np.ndarray, k: int = 10):
        """e4m3 coarse pass + exact bf16 re-rank (gallery created with Gallery.FP8)."""
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        s = np.zeros((nq, k), np.float32)
        i = np.zeros((nq, k), np.int64)
        self.ctx._check(lib().fr_gallery_search_fp8(self.h, q.ctypes.data, nq, k, FR_MEM_HOST, s.ctypes.data, i.ctypes.data))
        return s, i

    def search_fp8_dev(self, q_ptr: int, nq: int, k: int, out_s_ptr: int, out_i_ptr: int):
        self.ctx._check(lib().fr_gallery_search_fp8(self.h, q_ptr, nq, k, FR_MEM_DEVICE, out_s_ptr, out_i_ptr))

    def search_packed_fp8_dev(self, q_ptr: int, nq: int, k: int, out_rec_ptr: int):
        self.ctx._check(lib().fr_gallery_search_packed_fp8(self.h, q_ptr, nq, k, FR_MEM_DEVICE, out_rec_ptr))

    def search_packed(self, queries: np.ndarray, k: int = 10) -> np.ndarray:
        """Local top-k as packed records (uint64: low word fp32 score bits, high word global index)."""
        q = np.ascontiguousarray(queries, np.float32)
        rec = np.zeros((q.shape[0], k), np.uint64)
        self.ctx._check(lib().fr_gallery_search_packed(self.h, q.ctypes.data, q.shape[0], k, FR_MEM_HOST, rec.ctypes.data))
        return rec

    def search_packed_dev(self, q_ptr: int, nq: int, k: int, out_rec_ptr: int):
        self.ctx._check(lib().fr_gallery_search_packed(self.h, q_ptr, nq, k, FR_MEM_DEVICE, out_rec_ptr))


def topk_merge(ctx: Context, scores: np.ndarray, idx: np.ndarray, k: int):
    """scores/idx: [parts, nq, k] host arrays -> merged [nq, k]."""
    s = np.ascontiguousarray(scores, np.float32)
    i = np.ascontiguousarray(idx, np.int64)
    parts, nq, kk = s.shape
    assert kk == k
    os_ = np.zeros((nq, k), np.float32)
    oi = np.zeros((nq, k), np.int64)
    ctx._check(lib().fr_topk_merge(ctx.h, s.ctypes.data, i.ctypes.data, parts, nq, k, FR_MEM_HOST,
                                   os_.ctypes.data, oi.ctypes.data))
    return os_, oi


def topk_merge_packed(ctx: Context, records: np.ndarray, k: int):
    """records: [parts, nq, k] uint64 host array -> merged (scores [nq,k], idx [nq,k])."""
    r = np.ascontiguousarray(records, np.uint64)
    parts, nq, kk = r.shape
    assert kk == k
    os_ = np.zeros((nq, k), np.float32)
    oi = np.zeros((nq, k), np.int64)
    ctx._check(lib().fr_topk_merge_packed(ctx.h, r.ctypes.data, parts, nq, k, FR_MEM_HOST, os_.ctypes.data, oi.ctypes.data))
    return os_, oi


def topk_merge_packed_dev(ctx: Context, rec_ptr: int, parts: int, nq: int, k: int, os_ptr: int, oi_ptr: int):
    ctx._check(lib().fr_topk_merge_packed(ctx.h, rec_ptr, parts, nq, k, FR_MEM_DEVICE, os_ptr, oi_ptr))


def topk_merge_dev(ctx: Context, s_ptr: int, i_ptr: int, parts: int, nq: int, k: int, os_ptr: int, oi_ptr: int):
    ctx._check(lib().fr_topk_merge(ctx.h, s_ptr, i_ptr, parts, nq, k, FR_MEM_DEVICE, os_ptr, oi_ptr))
