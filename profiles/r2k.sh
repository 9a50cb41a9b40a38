#!/bin/bash
python -m pytest tests -m gpu -q -k "table_full" 2>&1 | tail -3
for w in c1 c2frame c4:1.0; do python profiles/kernels_of.py $w; done
MOT_ENV_NOTE="mode1" MOT_UF_MODE=1 python profiles/kernels_of.py c1
MOT_ENV_NOTE="mode1" MOT_UF_MODE=1 python profiles/kernels_of.py c2frame
MOT_ENV_NOTE="light1024" MOT_UF_LIGHT=1024 python profiles/kernels_of.py c4:1.0
MOT_ENV_NOTE="mode1" MOT_UF_MODE=1 TOPK=6 python profiles/kernels_of.py c4:1.0
