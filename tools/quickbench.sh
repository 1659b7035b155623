#!/bin/bash
# quickbench.sh "label" [bench args...]: one-line summary of a bench.py run (GPU box helper)
label=$1; shift
python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', d['kernel'], d['config']['kind'], round(d['value'],1), 'GB/s frac', round(d['roofline']['frac'],4), 'ratio', d['comp_ratio'], 'clk', d['clocks']['sm_mhz'])"
