#!/bin/bash
# multi-GPU session: N = number of visible GPUs.  2-GPU correctness tests (when N >= 2) + the bench at N ranks.
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "visible GPUs: $N"
if [ "$N" -ge 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --no-header -rf --timeout 800 > gpurun_out/r2_dist_tests_n$N.log 2>&1; tail -6 gpurun_out/r2_dist_tests_n$N.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); g=d['gallery_1toN']; print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], {k:g[k] for k in ('value','gallery_rows_total','ms_per_batch','merge')}, g.get('fp8'))"
tail -3 gpurun_out/r2_bench_n$N.err
