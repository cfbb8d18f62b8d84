#!/usr/bin/env python
"""prints the best variants of a tuner log (tools/_build/tune_* output)"""
import json
import sys

for path in sys.argv[1:]:
    rows = []
    for l in open(path):
        if l.startswith("{"):
            try:
                rows.append(json.loads(l))
            except Exception:
                pass
    ok = [r for r in rows if "ms_med" in r]
    ok.sort(key=lambda r: r["ms_med"])
    print("== %s: %d variants, %d errors" % (path, len(ok), len(rows) - len(ok)))
    ref = ok[0]["maxdiff"] if ok else 0
    for r in ok[:10]:
        print("  %-46s med %.4f best %.4f ms  %7.1f GB/s regs %3d cta/sm %2d thr %3d smem %6d diff %.2g" % (
            r["variant"], r["ms_med"], r["ms_best"], r["gbs_med"], r["regs"], r["ctas_per_sm"], r["threads"], r["smem"], r["maxdiff"]))
