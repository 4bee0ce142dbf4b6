"""Regenerates profiles/README.md and profiles/r1_step_summary.json from the raw captures in this folder.

    python profiles/make_summary.py

Inputs (all produced under gpurun on one B200, ncu passes only after the same command exited 0 without ncu):
  r1_bench.json                     python bench.py                                    (one JSON line)
  r1_bench_reference.json           python bench.py --impl reference
  r1_bench_n2.json / _n4 / _n8      torchrun ... bench.py --gpus N
  r1_launches_step.csv              ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv
                                    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery
  r1_tc_family_ncu_metrics.csv      ncu --metrics <dram, tensor pipe, L2 hit> -k regex:halo_gemm|shift_gemm (one step)
  r1_other_kernels_ncu_metrics.csv  ncu --metrics <duration, dram, issue active> on every other kernel (one step)
"""
import collections
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_ncu(name):
    rows = list(csv.reader(open(os.path.join(HERE, name))))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    out = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        out.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = v * scale      # durations in us, sizes in bytes
    return list(out.values())


def short(name):
    n = name.replace("<unnamed>::", "").replace("void ", "")
    return n.split("(")[0]


def main():
    bench = json.load(open(os.path.join(HERE, "r1_bench.json")))
    ref = json.load(open(os.path.join(HERE, "r1_bench_reference.json")))
    launches = load_ncu("r1_launches_step.csv")
    starts = [i for i, k in enumerate(launches) if "det_preprocess" in k["name"]]
    step = launches[starts[-2]:starts[-1]]
    fam = collections.OrderedDict()
    for k in step:
        n = short(k["name"])
        if "halo_gemm" in n or "shift_gemm" in n:
            n = "tc::halo_gemm / tc::shift_gemm (IResNet convs + FC, tcgen05 bf16)"
        f = fam.setdefault(n, {"us": 0.0, "launches": 0})
        f["us"] += k["gpu__time_duration.sum"]
        f["launches"] += 1
    total = sum(f["us"] for f in fam.values())
    for f in fam.values():
        f["share"] = f["us"] / total
    json.dump({"step_launches": len(step), "sum_us": total, "families": fam},
              open(os.path.join(HERE, "r1_step_summary.json"), "w"), indent=1)

    L = []
    A = L.append
    A("# Round 1 profiles (B200, `bench.py` workload: 64 frames 640x640 + 512 faces per step)\n")
    A("Regenerate this file with `python profiles/make_summary.py` (it only reads the raw files in this folder).")
    A("All captures: `gpurun`, one B200, after the same command exited 0 without ncu.  Per-launch times")
    A("under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
    A("## Bench line (`r1_bench.json`, `python bench.py`, defaults)\n")
    r, e, g, c = bench["roofline"], bench["e2e"], bench.get("gallery_1toN"), bench.get("cpu_baseline")
    A(f"* value {bench['value']:.0f} faces/s (device-resident inputs), e2e {e['value']:.0f} faces/s (pinned host buffers,")
    A(f"  H2D {e['h2d_bytes_per_step'] / 1e6:.1f} MB + D2H {e['d2h_bytes_per_step'] / 1e6:.2f} MB per step inside the timed region), "
      f"{bench['ms_per_step']:.2f} ms / step, {bench['gpu_launches'] // bench['steps']} launches / step")
    A(f"* roofline (tcgen05 conv family): {r['achieved']:.0f} TFLOP/s of {r['peak']} measured sustained bf16 = {r['frac']:.3f};")
    A(f"  share of step {r['share_of_step']:.2f}; DRAM traffic {r['traffic'] / 1e6:.0f} MB per launch (ncu)")
    if g:
        A(f"* 1:N search: {g['value']:.0f} queries/s on a {g['rows_per_gpu'] / 1e6:.2f} M-row shard = "
          f"{g['gemm_tflops_per_gpu']:.0f} TFLOP/s ({g['frac_of_sustained_bf16_peak']:.2f} of peak)")
    if c:
        A(f"* CPU baseline (oracle port, {c['cores']} intra-op threads like the reference): {c['value']:.1f} faces/s; "
          f"`--impl reference` (all {ref['cpu_baseline']['cores']} host threads): {ref['value']:.1f} faces/s (`r1_bench_reference.json`)")
    A(f"* stage times (ms / step): {bench['detail']['stage_ms_per_step']}")
    A(f"* clocks during the timed region: {bench['clocks']}\n")
    for n in (2, 4, 8):
        p = os.path.join(HERE, f"r1_bench_n{n}.json")
        if os.path.exists(p):
            b = json.load(open(p))
            gal = b.get("gallery_1toN") or {}
            A(f"* N={n} (`r1_bench_n{n}.json`, torchrun, `--steps 10 --warmup 3`): {b['value']:.0f} faces/s, e2e {b['e2e']['value']:.0f}"
              + (f"; 1:N {gal['value']:.0f} queries/s over {gal['gallery_rows_total'] / 1e6:.2f} M rows (row-sharded, NCCL all-gather + merge)" if gal else ""))
    A("")
    A("## Kernel shares of one step (`r1_launches_step.csv`: `ncu --metrics gpu__time_duration.sum`)\n")
    A(f"{len(step)} launches, {total:.0f} us summed.\n")
    A("| kernel family | launches / step | us / step | share |\n|---|---|---|---|")
    for n, f in sorted(fam.items(), key=lambda x: -x[1]["us"]):
        A(f"| {n} | {f['launches']} | {f['us']:.1f} | {100 * f['share']:.1f} % |")
    A("")
    A("## tcgen05 conv family, per launch (`r1_tc_family_ncu_metrics.csv`)\n")
    tc = load_ncu("r1_tc_family_ncu_metrics.csv")
    tsum = sum(k["gpu__time_duration.sum"] for k in tc)
    dsum = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in tc)
    A(f"Sum over the {len(tc)} launches of a step: {tsum:.0f} us, {dsum / 1e9:.2f} GB DRAM traffic "
      f"({dsum / len(tc) / 1e6:.0f} MB per launch).  Launches 16-41 are the 26 identical 256-channel 14x14 convs "
      "(conv1 / conv2 alternate).\n")
    A("| # | kernel | us | DRAM read MB | DRAM write MB | tensor pipe active % | L2 hit % |\n|---|---|---|---|---|---|---|")
    tp = next((m for m in tc[0] if "pipe_tensor" in m), None)
    l2 = next((m for m in tc[0] if "hit_rate" in m), None)
    for i, k in enumerate(tc):
        if 18 <= i <= 41:
            continue
        A(f"| {i} | {short(k['name']).replace('tc::', '')} | {k['gpu__time_duration.sum']:.1f} | {k['dram__bytes_read.sum'] / 1e6:.0f} | "
          f"{k['dram__bytes_write.sum'] / 1e6:.0f} | {k.get(tp, float('nan')):.1f} | {k.get(l2, float('nan')):.0f} |")
    A("")
    A("## SCRFD and the bandwidth / latency bound kernels, per launch (`r1_other_kernels_ncu_metrics.csv`)\n")
    A("| kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s | issue active % |\n|---|---|---|---|---|---|")
    for k in load_ncu("r1_other_kernels_ncu_metrics.csv"):
        us = k["gpu__time_duration.sum"]
        rd, wr = k["dram__bytes_read.sum"], k["dram__bytes_write.sum"]
        A(f"| {short(k['name'])} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | "
          f"{k['smsp__issue_active.avg.pct_of_peak_sustained_active']:.0f} |")
    A("")
    A(open(os.path.join(HERE, "notes.md")).read())
    open(os.path.join(HERE, "README.md"), "w").write("\n".join(L))
    print("wrote README.md,", len(step), "launches,", round(total), "us")


if __name__ == "__main__":
    main()
