"""Regenerates profiles/README.md and profiles/r2_step_summary.json from the raw captures in this folder
(the single generator of that file; round 1's summary is frozen in r1_README.md).

    python profiles/make_summary.py

Inputs (all produced under gpurun on one B200; ncu passes only after the same command exited 0 without ncu,
see dev/gpu_final_a.sh / dev/gpu_final_b.sh):
  r2_bench.json                  python bench.py                                    (one JSON line)
  r2_bench_reference.json        python bench.py --impl reference --steps 5 --warmup 1
  r2_bench_n2.json / _n4 / _n8   torchrun ... bench.py --gpus N   (when captured)
  r2_launches_step.csv           ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv
                                 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery
  r2_step_ncu_metrics.csv        ncu --metrics <duration, dram read/write, tensor pipe, issue active, L2 hit>
                                 -s 270 -c 200 of the same command (every kernel of at least one whole step)
  r2_halo_gemm2_full.txt         ncu --set full of tc::halo_gemm2_kernel (raw page excerpt; the .ncu-rep stays in gpurun_out/)
  r2_sass_opcodes.txt            python profiles/sass_opcodes.py
  r2_*_sweep.txt                 dev/sweep_env.py A/B runs quoted in DESIGN.md section 8
"""
import collections
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
TC_FLOP_PER_FACE = (12.6187 - 2 * 21.6760e-3) * 1e9     # conv + FC MACs x 2 minus the stem (bench.py)


def load_ncu(name):
    rows = list(csv.reader(open(os.path.join(HERE, name))))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    out = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        out.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = v * scale      # durations in us, sizes in bytes
    return list(out.values())


def short(name):
    n = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    return n.split("(")[0]


def one_step(launches):
    """The launches between two consecutive det_preprocess_kernel launches (one pipeline step)."""
    starts = [i for i, k in enumerate(launches) if "det_preprocess" in k["name"]]
    for a, b in zip(starts, starts[1:]):
        if b - a >= 90:        # a pipeline step (config loops launch shorter sequences)
            return launches[a:b]
    raise SystemExit("no complete step in the capture")


def is_tc(n):
    return "halo_gemm" in n or "shift_gemm" in n


def main():
    bench = json.load(open(os.path.join(HERE, "r2_bench.json")))
    ref = json.load(open(os.path.join(HERE, "r2_bench_reference.json")))
    step = one_step(load_ncu("r2_launches_step.csv"))
    fam = collections.OrderedDict()
    for k in step:
        n = short(k["name"])
        if is_tc(n):
            n = "tc::halo_gemm2 / halo_gemm / shift_gemm (IResNet convs + FC, tcgen05 bf16)"
        f = fam.setdefault(n, {"us": 0.0, "launches": 0})
        f["us"] += k["gpu__time_duration.sum"]
        f["launches"] += 1
    total = sum(f["us"] for f in fam.values())
    for f in fam.values():
        f["share"] = f["us"] / total
    met = one_step(load_ncu("r2_step_ncu_metrics.csv"))
    tc = [k for k in met if is_tc(short(k["name"]))]
    tp = next(m for m in tc[0] if "pipe_tensor" in m)
    l2 = next(m for m in tc[0] if "hit_rate" in m)
    tsum = sum(k["gpu__time_duration.sum"] for k in tc)
    dsum = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in tc)
    tp_w = sum(k[tp] * k["gpu__time_duration.sum"] for k in tc) / tsum
    json.dump({"step_launches": len(step), "sum_us": total, "families": fam,
               "tc_family": {"launches": len(tc), "sum_us": tsum, "dram_bytes": dsum, "traffic_bytes_per_launch": dsum / len(tc),
                             "time_weighted_tensor_pipe_pct": tp_w}},
              open(os.path.join(HERE, "r2_step_summary.json"), "w"), indent=1)
    json.dump({"traffic_bytes_per_launch": dsum / len(tc), "launches": len(tc),
               "source": "r2_step_ncu_metrics.csv: dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 conv family of one step"},
              open(os.path.join(HERE, "r2_tc_traffic.json"), "w"), indent=1)

    L = []
    A = L.append
    A("# Round 2 profiles (B200, `bench.py` workload: 64 frames 640x640 + 512 faces per step)\n")
    A("Regenerate this file with `python profiles/make_summary.py` (it only reads the raw files in this folder; round 1:")
    A("`r1_README.md`).  All captures: `gpurun`, one B200, ncu only after the same command exited 0 without it.  Per-launch")
    A("times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
    A("## Bench line (`r2_bench.json`, `python bench.py`, defaults: 20 steps, 5 warm-up)\n")
    r, e, g, c = bench["roofline"], bench["e2e"], bench.get("gallery_1toN"), bench.get("cpu_baseline")
    A(f"* value {bench['value']:.0f} faces/s (device-resident inputs), e2e {e['value']:.0f} faces/s (pinned host buffers,")
    A(f"  H2D {e['h2d_bytes_per_step'] / 1e6:.1f} MB + D2H {e['d2h_bytes_per_step'] / 1e6:.2f} MB per step inside the timed region), "
      f"{bench['ms_per_step']:.2f} ms / step, {bench['gpu_launches'] // bench['steps']} launches / step")
    A(f"* roofline (tcgen05 conv family): {r['achieved']:.0f} TFLOP/s of {r['peak']} measured sustained bf16 = {r['frac']:.3f};")
    A(f"  share of step {r['share_of_step']:.2f}")
    if g:
        A(f"* 1:N search, bf16: {g['value']:.0f} queries/s on a {g['rows_per_gpu'] / 1e6:.2f} M-row shard = "
          f"{g['gemm_tflops_per_gpu']:.0f} TFLOP/s ({g['frac_of_sustained_bf16_peak']:.2f} of peak; measured right after the 20-step pipeline run, on the power cap)")
        f8 = g.get("fp8") or {}
        if "value" in f8:
            A(f"* 1:N search, e4m3 coarse pass + exact bf16 re-rank: {f8['value']:.0f} queries/s ({f8['ms_per_batch']:.2f} ms per 4096 queries, "
              f"{f8['coarse_tflops_per_gpu']:.0f} TFLOP/s-equivalent), agreement with the bf16 top-10 {100 * f8['agreement_with_bf16_top10']:.3f} %")
    if c:
        A(f"* CPU baseline (oracle port, {c['cores']} intra-op threads like the reference): {c['value']:.1f} faces/s; "
          f"`--impl reference` (all {ref['cpu_baseline']['cores']} host threads): {ref['value']:.1f} faces/s (`r2_bench_reference.json`)")
    A(f"* stage times (ms / step): {json.dumps({k: round(v, 3) for k, v in bench['detail']['stage_ms_per_step'].items()})}")
    cf = bench["detail"].get("configs", {})
    if cf:
        c1, c2, c3 = cf["config1_compare_batch1"], cf["config2_det_only_batch64"], cf["config3_embed_only_batch1024"]
        A(f"* BASELINE configs measured on their own: (1) compare mode, batch 1, host buffers: {c1['ms_per_compare_median']:.2f} ms per compare "
          f"(CPU port, 4 threads: {c1.get('cpu_port_ms_per_compare_4threads', float('nan')):.1f} ms); (2) SCRFD only, batch 64: "
          f"{c2['ms_per_batch']:.2f} ms = {c2['frames_per_s']:.0f} frames/s; (3) ArcFace only, batch 1024: {c3['ms_per_batch']:.2f} ms = "
          f"{c3['faces_per_s']:.0f} faces/s ({c3['tflops']:.0f} TFLOP/s)")
    A(f"* clocks during the timed region: {bench['clocks']}\n")
    for n in (2, 4, 8):
        p = os.path.join(HERE, f"r2_bench_n{n}.json")
        if os.path.exists(p):
            b = json.load(open(p))
            gal = b.get("gallery_1toN") or {}
            f8 = gal.get("fp8") or {}
            A(f"* N={n} (`r2_bench_n{n}.json`, torchrun): {b['value']:.0f} faces/s, e2e {b['e2e']['value']:.0f}"
              + (f"; 1:N bf16 {gal['value']:.0f} queries/s over {gal['gallery_rows_total'] / 1e6:.2f} M rows (one all-gather of packed records + merge)" if gal else "")
              + (f", fp8 {f8['value']:.0f} queries/s" if "value" in f8 else ""))
    A("")
    A("## Kernel shares of one step (`r2_launches_step.csv`: `ncu --metrics gpu__time_duration.sum`)\n")
    A(f"{len(step)} launches, {total:.0f} us summed.\n")
    A("| kernel family | launches / step | us / step | share |\n|---|---|---|---|")
    for n, f in sorted(fam.items(), key=lambda x: -x[1]["us"]):
        A(f"| {n} | {f['launches']} | {f['us']:.1f} | {100 * f['share']:.1f} % |")
    A("")
    A("## tcgen05 conv family, per launch (`r2_step_ncu_metrics.csv`)\n")
    flop = 512 * TC_FLOP_PER_FACE
    A(f"Sum over the {len(tc)} launches of a step: {tsum:.0f} us ({flop / tsum / 1e6:.0f} TFLOP/s under ncu), {dsum / 1e9:.2f} GB DRAM traffic "
      f"({dsum / len(tc) / 1e6:.0f} MB per launch), time-weighted tensor-pipe activity {tp_w:.1f} %.  The 256 / 512-channel 3x3 stride-1 layers are "
      "`halo_gemm2_kernel` (CTA pairs); the run of identical 256-channel 14x14 launches is abbreviated.\n")
    A("| # | kernel | us | DRAM read MB | DRAM write MB | tensor pipe active % | L2 hit % |\n|---|---|---|---|---|---|---|")
    for i, k in enumerate(tc):
        if 18 <= i <= 41:
            continue
        A(f"| {i} | {short(k['name']).replace('tc::', '')} | {k['gpu__time_duration.sum']:.1f} | {k['dram__bytes_read.sum'] / 1e6:.0f} | "
          f"{k['dram__bytes_write.sum'] / 1e6:.0f} | {k.get(tp, float('nan')):.1f} | {k.get(l2, float('nan')):.0f} |")
    mid = tc[18:42]
    if mid:
        A(f"| 18-41 | {short(mid[0]['name']).replace('tc::', '')} x {len(mid)} | {sum(k['gpu__time_duration.sum'] for k in mid) / len(mid):.1f} avg | "
          f"{sum(k['dram__bytes_read.sum'] for k in mid) / len(mid) / 1e6:.0f} | {sum(k['dram__bytes_write.sum'] for k in mid) / len(mid) / 1e6:.0f} | "
          f"{sum(k[tp] for k in mid) / len(mid):.1f} | {sum(k[l2] for k in mid) / len(mid):.0f} |")
    A("")
    A("## SCRFD and the bandwidth / latency bound kernels, per launch (`r2_step_ncu_metrics.csv`)\n")
    A("| kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s | issue active % |\n|---|---|---|---|---|---|")
    ia = "smsp__issue_active.avg.pct_of_peak_sustained_active"
    for k in met:
        if is_tc(short(k["name"])):
            continue
        us = k["gpu__time_duration.sum"]
        rd, wr = k["dram__bytes_read.sum"], k["dram__bytes_write.sum"]
        A(f"| {short(k['name'])} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | {k.get(ia, float('nan')):.0f} |")
    A("")
    notes = os.path.join(HERE, "r2_notes.md")
    if os.path.exists(notes):
        A(open(notes).read())
    open(os.path.join(HERE, "README.md"), "w").write("\n".join(L))
    print("wrote README.md,", len(step), "launches,", round(total), "us; tc family", round(tsum), "us, tensor pipe", round(tp_w, 1))


if __name__ == "__main__":
    main()
