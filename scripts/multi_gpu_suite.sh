#!/bin/bash
# Every named workload of bench.py at N GPUs of one box (torchrun, one rank per GPU) + the bare PCIe probe.
#   scripts/multi_gpu_suite.sh N OUT_DIR [steps]
# Lines go to OUT_DIR/scale_N.jsonl (one JSON line per workload) and OUT_DIR/pcie_N.json.
N=${1:-2}; OUT=${2:-gpurun_out/r2}; STEPS=${3:-20}
mkdir -p "$OUT"; : > "$OUT/scale_$N.jsonl"
run() {
  if [ "$N" = 1 ]; then python "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; fi
}
# SUITE="workload:layout ..." restricts the list (default: every workload); NO_PCIE=1 skips the copy probe
SUITE=${SUITE:-"c3_4x4_etu:full c3_4x4_etu:compact c2_2x2_eva_dense:full c2_2x2_eva_dense_stats:full c4_sweep:full c5_mixed:full c5_mixed:compact c2_2x2_eva:full c1_siso_epa:full"}
for spec in $SUITE; do
  set -- ${spec%%:*} ${spec##*:}
  run bench.py --gpus "$N" --steps "$STEPS" --no-cpu-baseline --workload "$1" --layout "$2" --patterns 16 2> "$OUT/scale_${N}_$1_$2.err" | tail -1 >> "$OUT/scale_$N.jsonl"
done
[ -z "$NO_PCIE" ] && run scripts/pcie_d2h_probe.py 2> "$OUT/pcie_$N.err" | tail -1 > "$OUT/pcie_$N.json"
python - "$OUT/scale_$N.jsonl" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
        print(d["n_gpus"], d["config"]["workload"][:18], d["arm"]["layout"][:7], round(d["value"]), (d["roofline"].get("frac") or 0), d["e2e"].get("value"),
              (d["e2e"].get("roofline") or {}).get("peak"), (d.get("value_api") or {}).get("value"), d["stats_checksum"]["sha1"][:12], d["clocks"]["reasons"])
    except Exception as e:
        print("ERR", e, line[:200])
PY
[ -z "$NO_PCIE" ] && cat "$OUT/pcie_$N.json"; true
