#!/bin/bash
# One GPU call that produces the round's ncu evidence (each bench command is first run plain, then under ncu).
# Reports are reduced to CSV / JSON on the box (the .ncu-rep files with source are ~30 MB each; gpurun returns 64 MB).
#   scripts/profile_round.sh OUT_DIR
OUT=${1:-gpurun_out/r2p}; mkdir -p "$OUT"; TMP=$(mktemp -d)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --launches-per-step 1 --batch 2048 --e2e-batch 256 --e2e-steps 1"
$B > "$OUT/plain_full.json" 2> "$OUT/plain_full.err" || { tail -5 "$OUT/plain_full.err"; exit 1; }
# launch lists of the bench commands (kernel shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_full.csv" $B > "$OUT/ncu_l1.log" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_compact.csv" $B --layout compact > "$OUT/ncu_l2.log" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_dense.csv" $B --workload c2_2x2_eva_dense --batch 8192 > "$OUT/ncu_l3.log" 2>&1
cap() {   # name kernel-regex skip extra bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o "$TMP/$name" $B "$@" > "$OUT/ncu_$name.log" 2>&1
  ncu -i "$TMP/$name.ncu-rep" --page raw --csv > "$OUT/$name.raw.csv" 2>/dev/null
  python scripts/sass_histogram.py "$TMP/$name.ncu-rep" > "$OUT/$name.sass.json" 2>/dev/null
  rm -f "$TMP/$name.ncu-rep"
}
cap slot_full slot_kernel 4
cap slot_compact slot_kernel 4 --layout compact
cap slot_c5_dataset slot_kernel 4 --workload c5_mixed
cap slot_c4_stats slot_kernel 4 --workload c4_sweep
cap gemm_ta dense_tc_ta 3 --workload c2_2x2_eva_dense --batch 16384
cap k3_mode2 ls_interp 3 --workload c2_2x2_eva_dense --batch 16384
cap tap_gains tap_gains 3
python scripts/bench_kernels.py > "$OUT/kernels.json" 2> "$OUT/kernels.err" || tail -5 "$OUT/kernels.err"
rm -rf "$TMP"; du -sh "$OUT"; ls "$OUT"
