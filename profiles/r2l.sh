#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -6
for l in 64 256 1024; do MOT_ENV_NOTE="light$l" MOT_UF_LIGHT=$l TOPK=4 python profiles/kernels_of.py c4:1.0; done
for l in 256 1024; do MOT_ENV_NOTE="light$l" MOT_UF_LIGHT=$l TOPK=4 python profiles/kernels_of.py c4:0.3; done
TOPK=4 python profiles/exp_uf.py 16 4 -- "MOT_UF_LIGHT=64" "MOT_UF_LIGHT=256" "MOT_UF_LIGHT=1024"
TOPK=3 python profiles/kernels_of.py c1; TOPK=3 python profiles/kernels_of.py c2frame
