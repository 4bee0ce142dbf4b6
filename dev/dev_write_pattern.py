"""Developer micro-benchmark (not a test): write-only DRAM rate by store pattern (1 GiB, 512 B per thread).
    python tests/dev_write_pattern.py"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerecognizeonnx_b200 import capi

L = C.CDLL(os.path.join(os.path.dirname(capi.LIB_PATH), "libfr_dev.so"))   # developer micro-benchmarks live outside the product library
L.fr_debug_write_pattern.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
ctx = capi.Context(0, capi.Weights(capi.FR_MODEL_DET, None, 1), capi.Weights(capi.FR_MODEL_REC, None, 1))
for mode, name in ((0, "16 B per lane, 512-B lane stride (stem)"), (1, "32 B per lane pair"), (2, "128 B per 8 lanes"), (3, "fully coalesced")):
    ms, gbs = C.c_float(), C.c_double()
    rc = L.fr_debug_write_pattern(ctx.h, mode, 5, C.byref(ms), C.byref(gbs))
    print(f"{name:42s}: rc={rc} {ms.value:7.3f} ms  {gbs.value:7.0f} GB/s written", flush=True)
