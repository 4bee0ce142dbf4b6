"""Developer micro-benchmark (not a test): per-SM TMA streaming rate from DRAM.
    python tests/dev_tma_stream.py"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerecognizeonnx_b200 import capi

L = C.CDLL(os.path.join(os.path.dirname(capi.LIB_PATH), "libfr_dev.so"))   # developer micro-benchmarks live outside the product library
L.fr_debug_tma_stream.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
ctx = capi.Context(0, capi.Weights(capi.FR_MODEL_DET, None, 1), capi.Weights(capi.FR_MODEL_REC, None, 1))
rows = 8 * 1024 * 1024          # 1 GiB of 128-byte rows (>> 126 MB L2)
for mode, name in ((2, "TMA+16B/row stores"), (4, "TMA+32B/row stores"), (5, "TMA+64B/row stores"), (3, "TMA+128B/row stores")):
    for box, adv in ((248, 128),):
        for stages in (2, 3):
            if stages * box * 128 > 200 * 1024:
                continue
            ms, gbs = C.c_float(), C.c_double()
            rc = L.fr_debug_tma_stream(ctx.h, rows, box, adv, stages, mode, 3, C.byref(ms), C.byref(gbs))
            print(f"{name} box_rows={box:4d} advance={adv:4d} stages={stages}: rc={rc} {ms.value:8.3f} ms  {gbs.value:8.0f} GB/s loaded"
                  f"  ({gbs.value * adv / box:8.0f} GB/s unique; writes add {0 if mode < 2 else gbs.value * 128 / box:8.0f} GB/s)")
