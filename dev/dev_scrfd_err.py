import numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from facerecognizeonnx_b200 import capi
from oracle import nets
dw = capi.Weights(capi.FR_MODEL_DET, None, 1)
ctx = capi.Context(0, dw, capi.Weights(capi.FR_MODEL_REC, None, 1))
rng = np.random.default_rng(31)
n = 3
x = ((rng.integers(0, 256, (n, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
got = ctx.scrfd_forward(x)
wd = dw.to_dict()
ref, taps = nets.scrfd_forward(wd, torch.from_numpy(x), return_taps=True)
wd64 = {k: v.astype(np.float64) for k, v in wd.items()}
# float64 reference to separate "our error" from "torch fp32's own error"
import torch.nn.functional as F
def _t(w, name): return torch.from_numpy(np.ascontiguousarray(w[name]))
nets._t = lambda w, name: torch.from_numpy(np.ascontiguousarray(w[name]))
ref64, taps64 = nets.scrfd_forward(wd64, torch.from_numpy(x.astype(np.float64)), return_taps=True)
for ti, (t, t64) in enumerate(zip(taps, taps64)):
    t = t.numpy(); t64 = t64.numpy()
    g = ctx.scrfd_tap(ti, n, t.shape[1:])
    rng_ = np.abs(t64).max()
    print(f"tap {ti:2d} range {rng_:8.3f}  gpu-vs-f64 {np.abs(g - t64).max()/rng_:.2e}  torch32-vs-f64 {np.abs(t - t64).max()/rng_:.2e}  gpu-vs-torch32 {np.abs(g - t).max()/rng_:.2e}")
for i, (g, r, r64) in enumerate(zip(got, ref, ref64)):
    r = r.numpy(); r64 = r64.numpy()
    print(f"head {i} range {np.abs(r64).max():7.3f} gpu-vs-f64 {np.abs(g - r64).max():.2e} torch32-vs-f64 {np.abs(r - r64).max():.2e} gpu-vs-torch32 {np.abs(g-r).max():.2e}")
