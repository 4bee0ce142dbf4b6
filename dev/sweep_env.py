"""Developer tool (not a test): run bench.py under several environment settings and print the
per-stage times, to A/B the env-selectable variants of the kernels / schedules on one box.
    python dev/sweep_env.py "FR_REC_CHUNK=0 FR_DET_CHUNK=0" "FR_REC_CHUNK=29" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = os.environ.get("SWEEP_STEPS", "10")
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3",
                        "--no-cpu-baseline", "--no-gallery"], env=env, capture_output=True, text=True, cwd=ROOT)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        st = d["detail"]["stage_ms_per_step"]
        print(f"{spec:60s} step {d['ms_per_step']:.3f} ms  value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  "
              f"scrfd {st['scrfd']:.3f} stem {st['stem']:.3f} trunk {st['trunk']:.3f} align {st['align']:.3f} "
              f"pre {st['preprocess']:.3f} nms {st['decode_nms']:.3f}  launches {d['gpu_launches']}  "
              f"c2 {d['detail']['configs']['config2_det_only_batch64']['ms_per_batch']:.3f} "
              f"c3 {d['detail']['configs']['config3_embed_only_batch1024']['ms_per_batch']:.3f} "
              f"c1 {d['detail']['configs']['config1_compare_batch1']['ms_per_compare_median']:.3f} "
              f"steps(min/med/max) {'/'.join('%.2f' % v for v in d['detail']['step_ms_min_median_max'])} "
              f"clk {d['clocks']['sm_mhz']} slowest {d['detail']['slowest_step_index']} "
              f"submit max {d['detail']['host_submit_ms_max']:.1f} ms at {d['detail']['host_submit_slowest_index']}", flush=True)
    except Exception as e:
        print(spec, "FAILED", e, r.stderr[-800:], flush=True)
