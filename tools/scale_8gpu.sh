#!/bin/bash
# Round-2 scaling runs on an 8-GPU box:  gpurun --gpus 8 -- bash tools/scale_8gpu.sh [tag]
# Host topology first, then the driver's own command at 8 GPUs (headline + e2e + configs array), e2e variants
# (no CPU binding / staged copy pipeline), and the headline at 4, 2, 1 GPUs.  Lines land in gpurun_out/<tag>.jsonl.
set -u
tag=${1:-r2_scale8}
out=gpurun_out/$tag.jsonl
mkdir -p gpurun_out; : > $out; : > gpurun_out/$tag.err
bash tools/host_topology.sh > gpurun_out/${tag}_topo.txt 2>&1
run() {   # N, extra args...
    local n=$1; shift
    echo "== N=$n $*" >> gpurun_out/$tag.err
    if [ "$n" = 1 ]; then
        timeout 600 python bench.py --gpus 1 "$@" 2>>gpurun_out/$tag.err | tail -1 >> $out
    else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
            bench.py --gpus $n "$@" 2>>gpurun_out/$tag.err | tail -1 >> $out
    fi
}
run 8 --steps 20 --warmup 5
run 8 --steps 20 --warmup 5 --configs none --no-bind
run 8 --steps 20 --warmup 5 --configs none --e2e-zerocopy
for n in 4 2 1; do run $n --steps 20 --warmup 5 --configs none --no-cpu-baseline; done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print('BAD', l[:200]); continue
    e = d.get('e2e') or {}
    print(d['n_gpus'], '| %.3e /s | %.3f ms | frac %.3f | e2e %s | ceiling %s (%.0f GB/s) | %s' % (
        d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('frac') or float('nan'),
        ('%.3e' % e['value']) if e else '-', ('%.3e' % e['pcie_aggregate']['ceiling_configurations_per_s']) if e else '-',
        e['pcie_aggregate']['GBs'] if e else 0, (e.get('api') or '')[28:60] + ' / ' + (e.get('cpu_binding') or '')[:40]))
    for c in d.get('configs', []):
        if 'value' in c:
            r = c['roofline']
            print('     %-100s | %.3e /s %.3f ms %s frac %.3f %s' % (c['name'][:100], c['value'], c['ms_per_step'], r['bound'], r['frac'] or 0, c['clocks']['reasons']))
        else:
            print('     ', c)
PY
