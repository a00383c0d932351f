#!/bin/bash
# DRAM bytes and L2 hit rate of the slot kernel against the size of the pilot-pattern pool (ncu, one launch of 2048 slots).
OUT=${1:-gpurun_out/r2}; mkdir -p "$OUT"
for p in 1 64 1024; do for l in full compact; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sector_op_read_hit_rate.pct \
      --clock-control none -k regex:slot_kernel -s 6 -c 1 --csv --log-file "$OUT/pool_${p}_${l}.csv" \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --launches-per-step 1 --batch 2048 --e2e-batch 256 --e2e-steps 1 \
      --patterns $p --layout $l > /dev/null 2>&1
done; done
grep -h slot_kernel "$OUT"/pool_*.csv | cut -c1-40,200-
