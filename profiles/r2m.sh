#!/bin/bash
for s in 2 3 4 6 8; do python bench.py --steps 20 --warmup 3 --streams $s --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streams', d['streams_per_gpu'], 'value', d['value'], 'ms', d['ms_per_step'], d['timed_regions_ms'], 'e2e', d['e2e']['value'], 'lat', d['single_frame_latency_us'])"; done
