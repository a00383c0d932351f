#!/bin/bash
# One GPU call that produces the round's ncu evidence (each bench command is first run plain, then under ncu):
#   scripts/profile_round.sh OUT_DIR
OUT=${1:-gpurun_out/r2p}; mkdir -p "$OUT"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --launches-per-step 1 --batch 2048 --e2e-batch 256 --e2e-steps 1"
$B > "$OUT/plain_full.json" 2> "$OUT/plain_full.err" || { tail -5 "$OUT/plain_full.err"; exit 1; }
# launch list of the default bench command (kernel shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_full.csv" $B > "$OUT/ncu_l1.log" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_compact.csv" $B --layout compact > "$OUT/ncu_l2.log" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_dense.csv" $B --workload c2_2x2_eva_dense --batch 8192 > "$OUT/ncu_l3.log" 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:slot_kernel -s 4 -c 1 -o "$OUT/slot_full" $B > "$OUT/ncu_f1.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:slot_kernel -s 4 -c 1 -o "$OUT/slot_compact" $B --layout compact > "$OUT/ncu_f2.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:slot_kernel -s 4 -c 1 -o "$OUT/slot_c5_dataset" $B --workload c5_mixed > "$OUT/ncu_f3.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:slot_kernel -s 4 -c 1 -o "$OUT/slot_c4_stats" $B --workload c4_sweep > "$OUT/ncu_f4.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_tc_ta -s 3 -c 1 -o "$OUT/gemm_ta" $B --workload c2_2x2_eva_dense --batch 16384 > "$OUT/ncu_f5.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:ls_interp -s 3 -c 1 -o "$OUT/k3_mode2" $B --workload c2_2x2_eva_dense --batch 16384 > "$OUT/ncu_f6.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:tap_gains -s 3 -c 1 -o "$OUT/tap_gains" $B > "$OUT/ncu_f7.log" 2>&1
python scripts/bench_kernels.py > "$OUT/kernels.json" 2> "$OUT/kernels.err" || tail -5 "$OUT/kernels.err"
ls -la "$OUT" | head -40
