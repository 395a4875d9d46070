#!/bin/bash
# Fused solve + all-gather against NCCL, N GPUs of one box:  gpurun --gpus N -- bash tools/gather_ngpu.sh N
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/gather_err.log | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print('%-28s value %.3e  ms/step %.3f  collective: %s' % (' '.join(sys.argv[1:]) or 'no collective', d['value'], d['ms_per_step'], d['config']['collective'][:60]))
except Exception as e:
    print('FAILED', ' '.join(sys.argv[1:]), l[:300])
" "$@" || { echo "FAILED $@"; tail -5 gpurun_out/gather_err.log; }; }
if [ "$2" = "peers" ]; then   # peer-store paths with the destination list replicated (emulates more peers on a small box)
  run --allgather-fused ipc --gather-replicate 4 --gather-rowwise | tee -a gpurun_out/gather_${N}gpu_peers.txt
  run --allgather-fused ipc --gather-replicate 4 | tee -a gpurun_out/gather_${N}gpu_peers.txt
  run --allgather-fused ipc | tee -a gpurun_out/gather_${N}gpu_peers.txt
  run --allgather-fused multicast | tee -a gpurun_out/gather_${N}gpu_peers.txt
  tail -3 gpurun_out/gather_err.log
  exit 0
fi
run | tee -a gpurun_out/gather_${N}gpu.txt
run --allgather | tee -a gpurun_out/gather_${N}gpu.txt
run --allgather-fused ipc | tee -a gpurun_out/gather_${N}gpu.txt
[ "$N" -le 2 ] && run --allgather-fused symm | tee -a gpurun_out/gather_${N}gpu.txt
run --allgather-fused multicast | tee -a gpurun_out/gather_${N}gpu.txt
tail -3 gpurun_out/gather_err.log
