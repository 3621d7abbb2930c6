"""Summaries of the ncu CSVs tools/profile_round2.sh writes (run here, on the CPU box):
  python tools/ncu_summarise.py traffic  <conv_dram.csv>  <out.json>      mean DRAM bytes per conv launch (bench.py's roofline.traffic)
  python tools/ncu_summarise.py launches <launches.csv>                   per-kernel time and share of one frame step"""
import csv
import json
import re
import sys


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    return re.sub(r"\(.*", "", name.replace("void ", "").replace("nhvr::", ""))


if sys.argv[1] == "traffic":
    per = {}
    for r in rows(sys.argv[2]):
        d = per.setdefault(r["ID"], {})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    tot = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in per.values()]
    out = {"clips": 8, "precision": "strict", "mean_dram_bytes_per_launch": sum(tot) / len(tot), "launches": len(tot),
           "mean_tensor_pipe_active_pct": sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for d in per.values()) / len(per),
           "source": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the conv launches of one strict frame step, "
                     "tools/profile_round2.sh)" % sys.argv[2]}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(out)
else:
    agg = {}
    for r in rows(sys.argv[2]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v * 1e3 if r["Metric Unit"] in ("ms", "msecond") else v
        a = agg.setdefault(short(r["Kernel Name"]), [0.0, 0])
        a[0] += v; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print("%d launches, %.1f us (cold-cache, serialised)" % (sum(a[1] for a in agg.values()), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %-44s n=%3d  %9.1f us  %5.1f %%" % (k, a[1], a[0], 100.0 * a[0] / tot))
