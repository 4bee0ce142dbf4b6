"""Developer tool: per-step GPU time of the graph-replayed pipeline over many short trials, with and without
the nvidia-smi clock sampler running, to look for sporadic slow trials."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from facerecognizeonnx_b200 import capi
dev = torch.device("cuda", 0)
det_w = capi.Weights(capi.FR_MODEL_DET, None, 1); rec_w = capi.Weights(capi.FR_MODEL_REC, None, 1)
ctx = capi.Context(0, det_w, rec_w)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n_img, K, n_rot = 64, 8, 4
g = torch.Generator(device="cpu").manual_seed(100)
frames = [torch.randint(0, 256, (n_img, 640, 640, 3), dtype=torch.uint8, generator=g).to(dev) for _ in range(n_rot)]
pad = torch.from_numpy(bench.synth_pad_faces(np.random.default_rng(200), n_img, K).view(np.uint8).reshape(n_img * K, 60).copy()).to(dev)
o = [torch.empty((n_img * K, 60), dtype=torch.uint8, device=dev), torch.empty(n_img, dtype=torch.int32, device=dev),
     torch.empty((n_img * K, 512), dtype=torch.float32, device=dev), torch.empty(n_img * K, dtype=torch.int32, device=dev)]
fb = 640 * 640 * 3
def step(i):
    base = frames[i % n_rot].data_ptr()
    ctx.pipeline_dev([base + j * fb for j in range(n_img)], 640, 640, 640 * 3, K, pad.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr())
for i in range(12): step(i)
torch.cuda.synchronize()
def trial(steps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    evs[0].record(stream)
    for i in range(steps):
        step(i); evs[i + 1].record(stream)
    t_submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return sum(per) / steps, max(per), t_submit * 1e3
for label, smi in (("no sampler", False), ("with nvidia-smi -lms 100", True), ("no sampler", False)):
    proc = None
    if smi:
        proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.DEVNULL)
        time.sleep(0.25)
    res = [trial(10) for _ in range(12)]
    if proc: proc.terminate()
    print(label, " mean/max step ms, submit ms:", " | ".join(f"{a:.2f}/{b:.2f}/{c:.1f}" for a, b, c in res), flush=True)
