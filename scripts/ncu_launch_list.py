#!/usr/bin/env python3
"""ncu --csv launch list (gpu__time_duration.sum per launch) -> per-kernel totals / shares as JSON.
usage: ncu_launch_list.py <launches.csv> <out.json> "<command that was profiled>" """
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) > iV:
        v = float(r[iV].replace(",", ""))
        v = v / 1e3 if r[iU] in ("ns", "nsecond") else (v * 1e3 if r[iU] in ("ms", "msecond") else v)
        k = r[iK].split("(")[0]
        tot[k] += v
        cnt[k] += 1
total = sum(tot.values())
out = {"command": sys.argv[3] if len(sys.argv) > 3 else "",
       "note": "per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes",
       "kernels": [{"kernel": k, "launches": cnt[k], "total_us": round(v, 1), "avg_us": round(v / cnt[k], 2),
                    "share": round(v / total, 4)} for k, v in tot.most_common()]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k in out["kernels"]:
    print("%-70s n=%4d avg %9.2f us share %.4f" % (k["kernel"][-70:], k["launches"], k["avg_us"], k["share"]))
