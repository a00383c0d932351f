"""Opcode histogram (executed warp instructions, shared-memory wavefronts, stall samples) of one kernel from an ncu
report's SASS source page:  python scripts/sass_histogram.py report.ncu-rep > hist.json"""
import collections
import csv
import io
import json
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia, isrc, iw, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("L1 Wavefronts Shared"), hdr.index("Warp Stall Sampling (All Samples)")
byop, wf, stall, execs = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[rows.index(hdr) + 1:]:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    op = r[isrc].split()
    o = (op[1] if op[0].startswith("@") else op[0]).rstrip(";")
    byop[o] += n
    execs[n] += 1
    for c, i in ((wf, iw), (stall, ist)):
        try:
            c[o] += int(r[i])
        except ValueError:
            pass
tot = sum(byop.values())
hot = max((n for n in execs if n > 0), key=lambda n: n * execs[n])
print(json.dumps({"kernel": rows[0][1] if rows and len(rows[0]) > 1 else "", "warp_instructions": tot,
                  "hot_loop": {"executions_per_instruction": hot, "static_instructions": execs[hot]},
                  "by_opcode": {o: {"executed": n, "pct": round(100 * n / tot, 2), "smem_wavefronts": wf[o], "stall_samples": stall[o]}
                                for o, n in byop.most_common(40)}}, indent=1))
