"""Summarise an ncu --csv log of (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) per GEMM launch:
total bytes, per-template totals, and the per-launch rows of encoder layer 12 (qkv, o, wi, wo). Usage: traffic_summary.py file.csv [out.json]"""
import csv, json, sys


def load(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        key = r["ID"]
        d = per.setdefault(key, dict(name=r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            mul = dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(unit, 1)
            d["read" if "read" in r["Metric Name"] else "write"] = v * mul
        else:
            mul = dict(ns=1e-6, us=1e-3, ms=1, s=1e3).get(unit, 1e-6)
            d["ms"] = v * mul
    return [per[k] for k in sorted(per, key=lambda x: int(x))]


if __name__ == "__main__":
    rows = load(sys.argv[1])
    tot_r = sum(r.get("read", 0) for r in rows); tot_w = sum(r.get("write", 0) for r in rows); tot_ms = sum(r.get("ms", 0) for r in rows)
    by = {}
    for r in rows:
        t = r["name"].split("(")[0]
        b = by.setdefault(t, [0, 0.0, 0.0, 0.0])
        b[0] += 1; b[1] += r.get("read", 0); b[2] += r.get("write", 0); b[3] += r.get("ms", 0)
    out = dict(launches=len(rows), dram_read_bytes=tot_r, dram_write_bytes=tot_w, ms_under_ncu=tot_ms,
               dram_bytes_per_launch=(tot_r + tot_w) / max(len(rows), 1),
               by_template={k: dict(launches=v[0], read_gb=round(v[1] / 1e9, 2), write_gb=round(v[2] / 1e9, 2), ms=round(v[3], 2)) for k, v in by.items()})
    big = [r for r in rows if r.get("ms", 0) > 0.5]
    out["largest_launches_sample"] = [dict(name=r["name"].split("(")[0][-24:], read_gb=round(r.get("read", 0) / 1e9, 2), write_gb=round(r.get("write", 0) / 1e9, 2),
                                           ms=round(r.get("ms", 0), 3)) for r in big[100:108]]
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)
