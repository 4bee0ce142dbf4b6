"""Developer micro-benchmark (not a test): DRAM rate of 2-D tiled TMA reads (+ epilogue-like writes)
over a [64*320 rows][320 px * 64 B] fp32 NHWC map, for several tile shapes.
    python tests/dev_tma_tiles.py"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerecognizeonnx_b200 import capi

L = C.CDLL(os.path.join(os.path.dirname(capi.LIB_PATH), "libfr_dev.so"))   # developer micro-benchmarks live outside the product library
L.fr_debug_tma_tiles.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_void_p, C.c_void_p]
ctx = capi.Context(0, capi.Weights(capi.FR_MODEL_DET, None, 1), capi.Weights(capi.FR_MODEL_REC, None, 1))
rows, pitch8 = 64 * 320, 320 * 8
shapes = (
    ("8x16 px tile, box 10x18", 144, 10, 1, 128, 8, 8),
    ("8x16 px tile, no halo", 128, 8, 1, 128, 8, 8),
    ("4x32 px tile, box 6x32", 256, 6, 1, 256, 4, 8),
    ("1x128 px strip, box 3x128", 256, 3, 4, 1024, 1, 8),
    ("2x320 px rows, box 4x320", 256, 4, 10, 2560, 2, 2),
    ("4x320 px rows, no halo", 256, 4, 10, 2560, 4, 2),
    ("1x320 px row, no halo", 256, 1, 10, 2560, 1, 8),
)
for name, bw8, bh, nx, adv_w8, adv_h, stages in shapes:
    for wmode, wname in ((0, "read only"), (1, "+16B/px-lane stores"), (2, "+coalesced stores")):
        ms, gbs = C.c_float(), C.c_double()
        rc = L.fr_debug_tma_tiles(ctx.h, rows, pitch8, bw8, bh, nx, adv_w8, adv_h, stages, wmode, 3, C.byref(ms), C.byref(gbs))
        print(f"{name:28s} stages={stages} {wname:22s}: rc={rc} {ms.value:7.3f} ms  {gbs.value:7.0f} GB/s unique read"
              f"{'' if wmode == 0 else ' (+ same written)'}", flush=True)
