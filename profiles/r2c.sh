#!/bin/bash
PK=multiple-object-tracking-lidar_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c_pytest.log
python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=1" "MOT_UF_MODE=2" "MOT_UF_MODE=2 MOT_UF_XBLOCKS=6" "MOT_UF_MODE=2 MOT_UF_XBLOCKS=2" "MOT_UF_MODE=2 MOT_UF_LIGHT=256" > gpurun_out/r2c_exp.log 2>&1
MOT_B200_LIB=$PWD/$PK/libmot_b200_stats1.so python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2c_stats1.log 2>&1
python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_uf_cross|k_cell_local" -s 6 -c 2 -o gpurun_out/r2c_uf python profiles/exp_uf.py 16 1 -- "MOT_UF_MODE=2" > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_pytest.log gpurun_out/r2c_exp.log; head -3 gpurun_out/r2c_stats1.log
tail -3 gpurun_out/r2c_ncu.log
