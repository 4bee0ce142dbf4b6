"""Developer tool: how many candidates reach NMS and how many survive on the bench's synthetic frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facerecognizeonnx_b200 import capi
det_w = capi.Weights(capi.FR_MODEL_DET, None, 1)
ctx = capi.Context(0, det_w, None)
g = torch.Generator(device="cpu").manual_seed(100)
frames = torch.randint(0, 256, (8, 640, 640, 3), dtype=torch.uint8, generator=g).numpy()
chw, scale = ctx.det_preprocess([f for f in frames])
heads = ctx.scrfd_forward(chw)
sc = np.concatenate([heads[0], heads[1], heads[2]], axis=1)[..., 0]
print("candidates > 0.5 per frame:", (sc > 0.5).sum(1).tolist())
kept = ctx.detect_batch([f for f in frames], 0.5, 0.4, cap=16384)
print("kept per frame:", [len(k) for k in kept])
