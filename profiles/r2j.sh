#!/bin/bash
python -m pytest tests -m gpu -q -k "overflow or table_full" 2>&1 | tail -5 > gpurun_out/r2j_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?" >> gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_pytest.log; tail -5 gpurun_out/r2j_bench.err; cat gpurun_out/r2j_bench.json | head -c 9000
