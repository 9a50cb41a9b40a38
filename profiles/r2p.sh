#!/bin/bash
MOT_UF_MODE=2 TOPK=6 python profiles/kernels_of.py c2frame
MOT_UF_MODE=2 TOPK=4 python profiles/kernels_of.py c1
TOPK=3 python profiles/exp_uf.py 16 4 -- "MOT_UF_MODE=2"
MOT_UF_MODE=2 python profiles/kernels_of.py c2frame > gpurun_out/r2p_plain.log 2>&1 && \
MOT_UF_MODE=2 ncu --set full --clock-control none --import-source on -k regex:"k_uf_fused|k_cell_local" -s 8 -c 2 -o gpurun_out/r2p_frame python profiles/kernels_of.py c2frame > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
