#!/bin/bash
# ncu evidence for the array-free dense statistics pipeline (bench workload c2_2x2_eva_dense_stats): launch list + one
# --set full capture each of the first pass (slot2 kernel with pilot vectors out) and of the scoring pass.
#   scripts/profile_dense_stats.sh OUT_DIR
OUT=${1:-gpurun_out/r2d}; mkdir -p "$OUT"
W="--workload c2_2x2_eva_dense_stats --batch 8192"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --launches-per-step 1"
$B $W > "$OUT/plain_dense_stats.json" 2> "$OUT/plain_dense_stats.err" || { tail -5 "$OUT/plain_dense_stats.err"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_dense_stats.csv" $B $W > "$OUT/ncu_l.log" 2>&1
# per pass the workload launches slot2_kernel twice: even launches = first pass, odd = scoring pass
scripts/ncu_one.sh dense_pass1 slot2_kernel 8 "$OUT" $W
scripts/ncu_one.sh dense_score slot2_kernel 9 "$OUT" $W
du -sh "$OUT"; ls "$OUT"
