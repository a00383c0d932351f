#!/bin/bash
# A/B of the plan staging: bulk-copied row ring (B2C_PLAN_BULK=1) against per-thread cp.async (B2C_PLAN_BULK=0)
#   scripts/ab_plan_bulk.sh OUT_DIR "workload layout" ...
OUT=${1:-gpurun_out/r2b}; shift; mkdir -p "$OUT"
[ $# -eq 0 ] && set -- "c3_4x4_etu full" "c3_4x4_etu compact" "c5_mixed full" "c2_2x2_eva full"
for cfg in "$@"; do
  for e in 1 0; do
    set -- $cfg
    f="$OUT/ab_${e}_$1_$2.json"
    B2C_PLAN_BULK=$e timeout 200 python bench.py --workload $1 --layout $2 --no-cpu-baseline --no-api --e2e-steps 1 --e2e-batch 256 > "$f" 2>/dev/null
    python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
t = d["roofline"].get("tensor") or {}
print(sys.argv[1], round(d["value"]), d["roofline"]["kernel_ms"], d["roofline"].get("frac"), t.get("score_ms"), d["stats_checksum"]["sha1"][:12], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  done
done
