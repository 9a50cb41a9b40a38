#!/bin/bash
nvidia-smi -L; nproc
python -m pytest tests -m gpu -q -k "batch_run" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r02_final_bench_2gpu.json 2> gpurun_out/r02_2gpu.err; echo "rc=$?"
tail -3 gpurun_out/r02_2gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_final_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','n_gpus','ms_per_step','timed_regions_ms','streams_per_gpu','host_wait']}); print(d['e2e'])"
