#!/bin/bash
# round-2 experiment b: hooking order, counters, scatter tile size, ncu of the two new kernels
PK=multiple-object-tracking-lidar_b200
python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=1" "MOT_UF_MODE=2" "MOT_UF_MODE=2 MOT_UF_SPLIT=1" > gpurun_out/r2b_hook1.log 2>&1
MOT_B200_LIB=$PWD/$PK/libmot_b200_stats1.so python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2b_stats1.log 2>&1
MOT_B200_LIB=$PWD/$PK/libmot_b200_stats0.so python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2b_stats0.log 2>&1
TOPK=5 MOT_B200_LIB=$PWD/$PK/libmot_b200_it16.so python profiles/exp_uf.py 16 3 -- "MOT_UF_MODE=1" > gpurun_out/r2b_it16.log 2>&1
TOPK=5 MOT_B200_LIB=$PWD/$PK/libmot_b200_it32.so python profiles/exp_uf.py 16 3 -- "MOT_UF_MODE=1" > gpurun_out/r2b_it32.log 2>&1
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_cross|k_cell_local" -s 6 -c 2 -o gpurun_out/r2b_uf python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2b_ncu.log 2>&1
for f in hook1 stats1 stats0 it16 it32; do echo "##### $f"; cat gpurun_out/r2b_$f.log; done
tail -5 gpurun_out/r2b_ncu.log
