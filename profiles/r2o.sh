#!/bin/bash
for fb in 5 6 8 10; do MOT_UF_FBLOCKS=$fb python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('fblocks $fb streams', d['streams_per_gpu'], 'value', d['value'], 'ms', d['ms_per_step'], d['timed_regions_ms'])"; done
MOT_UF_PRIO=0 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('prio0', 'value', d['value'], 'ms', d['ms_per_step'], d['timed_regions_ms'])"
