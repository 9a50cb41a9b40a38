#!/bin/bash
PK=multiple-object-tracking-lidar_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_pytest.log
TOPK=7 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=1" "MOT_UF_MODE=2" "MOT_UF_FBLOCKS=4" "MOT_UF_FBLOCKS=12" "MOT_UF_XMODE=1 MOT_UF_PHASES=7,24" > gpurun_out/r2g_exp.log 2>&1
for mb in 4 8; do TOPK=3 MOT_B200_LIB=$PWD/$PK/libmot_b200_f$mb.so python profiles/exp_uf.py 16 4 -- "MOT_UF_FBLOCKS=$mb" "MOT_UF_FBLOCKS=$((2*mb))" > gpurun_out/r2g_f$mb.log 2>&1; done
MOT_B200_LIB=$PWD/$PK/libmot_b200_stats1.so python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2g_stats1.log 2>&1
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_fused" -s 3 -c 1 -o gpurun_out/r2g_uf python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2g_ncu.log 2>&1
cat gpurun_out/r2g_pytest.log gpurun_out/r2g_exp.log; grep -A4 "===" gpurun_out/r2g_f4.log gpurun_out/r2g_f8.log; head -3 gpurun_out/r2g_stats1.log
tail -3 gpurun_out/r2g_ncu.log
