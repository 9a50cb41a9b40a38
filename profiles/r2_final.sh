#!/bin/bash
# round-2 evidence: default bench line, reference arm, ncu launch list under the bench's conditions, full captures
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_final_bench_reference.json 2>> gpurun_out/r02_final_bench.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --streams 1 --reps 1 --no-configs --no-cpu-baseline > gpurun_out/r02_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_ncu_launches.csv python bench.py --steps 2 --warmup 3 --streams 1 --reps 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_fused|k_cell_local|k_rs_scatter|k_cells_write" -s 24 -c 6 -o gpurun_out/r02_final_full python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
# the small-frame path (c1 frame): latency table, then a full capture of its five kernels as plain launches
python profiles/latency.py 60 > gpurun_out/r02_latency.txt 2>&1; python profiles/host_call.py 60 >> gpurun_out/r02_latency.txt 2>&1
MOT_SMALL_GRAPH=0 python profiles/small_frame_once.py 3 > gpurun_out/r02_plain3.log 2>&1 && \
MOT_SMALL_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_fs_" -s 10 -c 5 -f -o gpurun_out/r02_final_small python profiles/small_frame_once.py 3 > gpurun_out/r02_ncu_small.log 2>&1
echo "ncu small rc=$?"
head -c 1500 gpurun_out/r02_final_bench.json; echo; cat gpurun_out/r02_final_bench_reference.json | head -c 1200; echo; tail -3 gpurun_out/r02_final_bench.err
