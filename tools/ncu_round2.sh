#!/bin/bash
# ncu --set full captures of the round-2 kernels, one gpurun call per group (gpurun_out/ is limited to 64 MiB):
#   bash tools/ncu_round2.sh tip ctr_sec2_tub3 ctr_sec3_tub3     tip kernel of the named robots, with source
#   bash tools/ncu_round2.sh bb                                  backbone kernel (ctr_sec3_tub4, 200 k configurations)
#   bash tools/ncu_round2.sh f32                                 packed FP32 kernel (ctr_sec3_tub5)
# every capture only after the same command line has exited 0 without ncu
set -x
B="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --configs none"
NCU="ncu --set full --clock-control none --import-source ${SRC:-off} -s 3 -c 1 -f"
what=$1; shift
case $what in
tip)
  for r in "$@"; do
    python bench.py --robot $r.xml $B > gpurun_out/r2_ncu_$r.plain.log 2>&1 || exit 1
    $NCU -k regex:fk_tip_kernel -o gpurun_out/r2_tip_$r python bench.py --robot $r.xml $B > gpurun_out/r2_ncu_$r.log 2>&1
  done ;;
bb)
  python bench.py --robot ctr_sec3_tub4.xml --backbone --batch 200000 $B > gpurun_out/r2_ncu_bb.plain.log 2>&1 || exit 1
  $NCU -k regex:backbone -o gpurun_out/r2_bb_sec3_tub4 python bench.py --robot ctr_sec3_tub4.xml --backbone --batch 200000 $B > gpurun_out/r2_ncu_bb.log 2>&1 ;;
f32)
  python bench.py --robot ctr_sec3_tub5.xml --f32 $B > gpurun_out/r2_ncu_f32.plain.log 2>&1 || exit 1
  $NCU -k regex:f32x2 -o gpurun_out/r2_f32x2_sec3_tub5 python bench.py --robot ctr_sec3_tub5.xml --f32 $B > gpurun_out/r2_ncu_f32.log 2>&1 ;;
esac
du -sh gpurun_out/
