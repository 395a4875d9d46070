#!/bin/bash
# Scaling runs of DESIGN.md §7: headline at 1/2/4/8 GPUs, then BASELINE configs 3-5 at 8 GPUs.
# Run on an 8-GPU box:  gpurun --gpus 8 -- bash tools/scale_8gpu.sh   (lines land in gpurun_out/scale8.jsonl)
set -u
out=gpurun_out/scale8.jsonl
mkdir -p gpurun_out; : > $out
run() {   # N, extra args...
    local n=$1; shift
    if [ "$n" = 1 ]; then
        timeout 600 python bench.py --gpus 1 "$@" 2>>gpurun_out/scale8.err | tail -1 >> $out
    else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
            bench.py --gpus $n "$@" 2>>gpurun_out/scale8.err | tail -1 >> $out
    fi
}
for n in 1 2 4 8; do run $n --no-cpu-baseline; done
run 8 --no-cpu-baseline --no-e2e --allgather
run 8 --no-cpu-baseline --no-e2e --robot ctr_sec3_tub3.xml --batch 2000000
run 8 --no-cpu-baseline --backbone --robot ctr_sec3_tub4.xml --batch 1000000
run 8 --no-cpu-baseline --no-e2e --robot ctr_sec3_tub5.xml --batch 12500000
run 8 --no-cpu-baseline --f32 --robot ctr_sec3_tub5.xml --batch 12500000
python - <<'PY'
import json
for l in open('gpurun_out/scale8.jsonl'):
    try: d = json.loads(l)
    except Exception: print('BAD', l[:200]); continue
    print(d['n_gpus'], d['config']['workload'][:70], '| %.3e %s | %.2f ms | frac %.3f | e2e %s' % (
        d['value'], d['unit'], d['ms_per_step'], (d.get('roofline') or {}).get('frac') or float('nan'),
        ('%.3e' % d['e2e']['value']) if d.get('e2e') else '-'))
PY
