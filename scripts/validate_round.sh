#!/bin/bash
# Round-end validation on one GPU: full GPU test suite, smoke(), the default bench line, the reference arm, one ncu capture
# of the headline kernel.   scripts/validate_round.sh OUT_DIR
OUT=${1:-gpurun_out/r2v}; mkdir -p "$OUT"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > "$OUT/tests_gpu.log" 2>&1; tail -4 "$OUT/tests_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log"
( time timeout 600 python bench.py ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; tail -3 "$OUT/bench_default.err"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; tail -3 "$OUT/bench_reference.err"
scripts/ncu_one.sh slot_full_bulk slot_kernel 4 "$OUT"
scripts/ncu_one.sh slot_compact_bulk slot_kernel 4 "$OUT" --layout compact
python - "$OUT" <<'PY'
import json, sys
o = sys.argv[1]
for f in ("bench_default.json", "bench_reference.json"):
    try:
        d = json.loads([l for l in open(f"{o}/{f}").read().splitlines() if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, d.get("value"), d.get("unit"), "frac", r.get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
              (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "unreadable", e)
PY
