#!/bin/bash
# final round-2 measurement, part A (no profiler): all GPU tests, the bench line, the reference arm
set -x
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/r2_gputests.log 2>&1; tail -6 gpurun_out/r2_gputests.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 400 gpurun_out/r2_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
