#!/bin/bash
# Dense-pipeline checks after a GEMM change: parity tests, then the c2 dense bench at two batch sizes (+ opt-in cluster form).
OUT=${1:-gpurun_out/r2}; mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense" 2>&1 | tail -3
for b in 4144 16384; do
  timeout 200 python bench.py --steps 20 --no-cpu-baseline --workload c2_2x2_eva_dense --batch $b > "$OUT/dense_$b.json" 2> "$OUT/dense_$b.err" || tail -5 "$OUT/dense_$b.err"
done
B2C_DENSE_CLUSTER=1 timeout 200 python bench.py --steps 20 --no-cpu-baseline --workload c2_2x2_eva_dense --batch 16384 > "$OUT/dense_16384_cluster.json" 2>/dev/null
python - "$OUT" <<'PY'
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + "/dense_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); t = d["roofline"]["tensor"]
        print(f.split("/")[-1], round(d["value"]), d["clocks"]["sm_mhz"], round(t["gemm_ms"], 4), round(t["useful_tflops"], 1), round(t["frac"], 3),
              round(t["frac_of_measured_tf32_matmul"], 3), round(t["gemm_share_of_step"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
