#!/bin/bash
# Host-pointer path (ctr_fk_batch_host): pipeline depth x chunk size sweep; prints the e2e rate of bench.py.
#   gpurun -- bash tools/tune_host_pipeline.sh
for s in 3 4 6; do for w in 1 2; do
  CTRFK_HOST_STREAMS=$s CTRFK_HOST_CHUNK_WAVES=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams $s waves $w e2e %.3e value %.3e' % (d['e2e']['value'], d['value']))"
done; done
