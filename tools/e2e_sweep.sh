#!/bin/bash
# e2e forms of ctr_fk_batch_host on one GPU: staged pipeline (default) with its stream / chunk settings, zero-copy form
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --configs none "$@" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('%.4e e2e  kernel %.3f ms  frac_ceiling %.3f  agg %.1f GB/s  api: %s' % (e['value'], d['ms_per_step'], e['frac_of_pcie_ceiling'], e['pcie_aggregate']['GBs'], e['api'][18:60]))
"; }
echo 'staged pipeline (default)'; run
echo 'zero-copy reads and writes'; run --e2e-zerocopy
for s in 3 6; do for w in 1 2; do echo "staged, streams $s waves $w"; CTRFK_HOST_STREAMS=$s CTRFK_HOST_CHUNK_WAVES=$w run; done; done
for d in 2 3 4; do for s in 4 6 8; do echo "staged, chunk = wave/$d, streams $s"; CTRFK_HOST_CHUNK_DIV=$d CTRFK_HOST_STREAMS=$s run; done; done
