#!/bin/bash
# Backbone benchmark matrix (AoS): fixed N / step mode, several MaxSamples.  Prints ms per step.
run() { python bench.py --backbone --robot ctr_sec3_tub4.xml --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-60s %8.3f ms  %.3e /s' % ('$*', d['ms_per_step'], d['value']))"; }
run --batch 200000
run --batch 1000000
run --batch 200000 --mode S
run --batch 800000 --max-samples 64 --nsamples 64
run --batch 800000 --max-samples 64 --nsamples 64 --mode S
run --batch 25000 --max-samples 2048 --nsamples 2048
run --batch 200000 --robot ctr_sec2_tub3.xml
run --batch 200000 --robot ctr_sec3_tub5.xml
