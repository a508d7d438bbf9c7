"""Condense ncu --csv logs.
  launches <ncu.csv> <out.csv> [header comment]   per-launch list (second half = the second forward): idx,kernel,grid,block,duration_us
  traffic  <ncu.csv> <out.json> <command string>  per-kernel-name totals of dram bytes and duration + bytes per conv launch"""
import csv, json, sys
from collections import defaultdict


def read(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    col = {n: h.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value")}
    recs = {}
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        i = int(r[col["ID"]])
        rec = recs.setdefault(i, {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        u = r[col["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        rec[r[col["Metric Name"]]] = v * scale
    return [recs[k] for k in sorted(recs)]


def short(name):
    n = name.split("(")[0].replace("void ", "").strip()
    head, _, tail = n.partition("<")          # keep template arguments, drop namespaces
    head = head.split("::")[-1]
    if head.startswith("unnamed>"):
        head = head[len("unnamed>"):]
    n = head + (("<" + tail) if tail else "")
    return n.replace("unnamed>::", "")


if sys.argv[1] == "launches":
    recs = read(sys.argv[2])
    starts = [i for i, r in enumerate(recs) if "pack_miso" in r["kernel"]]
    recs = recs[starts[-1]:] if starts else recs[len(recs) // 2:]   # the last forward
    with open(sys.argv[3], "w") as f:
        if len(sys.argv) > 4:
            f.write("# " + sys.argv[4] + "\n")
        f.write("idx,kernel,grid,block,duration_us\n")
        for i, r in enumerate(recs):
            f.write(f"{i},{short(r['kernel'])},{r['grid'].replace(',', ' ')},{r['block'].replace(',', ' ')},{r['gpu__time_duration.sum']:.2f}\n")
    tot = defaultdict(lambda: [0, 0.0])
    for r in recs:
        tot[short(r["kernel"])][0] += 1
        tot[short(r["kernel"])][1] += r["gpu__time_duration.sum"]
    all_us = sum(v[1] for v in tot.values())
    for k, (n, us) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print(f"{k:36s} n={n:4d} {us / 1e3:8.3f} ms {100 * us / all_us:5.1f} %")
else:
    recs = read(sys.argv[2])
    recs = recs[len(recs) // 2:]
    out = {"command": sys.argv[4]}
    agg = defaultdict(lambda: {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "duration_us": 0.0})
    for r in recs:
        a = agg[short(r["kernel"]).split("<")[0]]
        a["launches"] += 1
        a["dram_read_bytes"] += r.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += r.get("dram__bytes_write.sum", 0.0)
        a["duration_us"] += r["gpu__time_duration.sum"]
    out.update(agg)
    main = max((k for k in agg if "prep" not in k), key=lambda k: agg[k]["duration_us"])
    tot = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in agg.values())
    out["family_dram_bytes_per_launch"] = tot / max(agg[main]["launches"], 1)
    import hashlib, os
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "misonet_b200", "csrc", "conv_rs.cu")
    out["kernel_source_sha1"] = hashlib.sha1(open(src, "rb").read()).hexdigest()   # bench.py reports a stale capture
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out)[:600])
