#!/bin/bash
# One --set full capture reduced to CSV on the box:  scripts/ncu_one.sh NAME KERNEL_REGEX SKIP OUT_DIR [bench args...]
NAME=$1; RX=$2; SKIP=$3; OUT=$4; shift 4; mkdir -p "$OUT"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --launches-per-step 1 --batch 2048 --e2e-batch 256 --e2e-steps 1"
$B "$@" > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o /tmp/$NAME $B "$@" > "$OUT/ncu_$NAME.log" 2>&1
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > "$OUT/$NAME.raw.csv" 2>/dev/null
python scripts/sass_histogram.py /tmp/$NAME.ncu-rep > "$OUT/$NAME.sass.json" 2>/dev/null
ls -la "$OUT/$NAME.raw.csv"
