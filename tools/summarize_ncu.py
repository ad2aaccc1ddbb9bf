"""Turn ncu CSV logs into the small tracked summaries under profiles/.

  launch list  (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv <cmd>):
      python tools/summarize_ncu.py launches L.csv "<cmd>" > profiles/launches_rNN_summary.tsv
  DRAM traffic (ncu --set full ... -k regex:conv_halo_kernel --csv --page raw --log-file R.csv <cmd>, or
                ncu -i X.ncu-rep --page raw --csv > R.csv):
      python tools/summarize_ncu.py traffic R.csv "<cmd>" > profiles/traffic_rNN_halo.json

The traffic JSON is what bench.py reads for `roofline.traffic` (bytes per launch of the halo kernels, averaged
over the five layers like `roofline.achieved`).
"""
from __future__ import annotations

import csv
import json
import re
import sys
from collections import OrderedDict


def _rows(path):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.strip() and not ln.startswith("==")]
    return list(csv.DictReader(lines))


def _short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name).strip()                  # drop the argument list
    return name.replace("void ", "void ", 1)


def launches(path: str, cmd: str) -> None:
    tot, cnt = OrderedDict(), OrderedDict()
    for r in _rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = _short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        tot[k] = tot.get(k, 0.0) + ns
        cnt[k] = cnt.get(k, 0) + 1
    mine = {k: v for k, v in tot.items() if "cfr::" in k}
    total = sum(mine.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}")
    print("kernel\tlaunches\ttotal_ms\tshare")
    for k, v in sorted(mine.items(), key=lambda kv: -kv[1]):
        print(f"{k}\t{cnt[k]}\t{v / 1e6:.3f}\t{v / total:.3f}")


def traffic(path: str, cmd: str) -> None:
    rows = _rows(path)
    out = []
    if rows and "Metric Name" in rows[0]:                        # long format (--csv of a live run)
        per = OrderedDict()
        for r in rows:
            per.setdefault(r["ID"], {"kernel": _short(r["Kernel Name"])})[r["Metric Name"]] = (
                float(r["Metric Value"].replace(",", "")), r.get("Metric Unit", ""))
        recs = per.values()
    else:                                                        # wide format (--page raw --csv of a report)
        units = rows[0]
        recs = []
        for r in rows[1:]:
            d = {"kernel": _short(r["Kernel Name"])}
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                if m in r:
                    d[m] = (float(r[m].replace(",", "")), units.get(m, ""))
            recs.append(d)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
             "msecond": 1.0, "nsecond": 1e-6}
    for d in recs:
        if "dram__bytes_read.sum" not in d:
            continue
        rd, ru = d["dram__bytes_read.sum"]
        wr, wu = d["dram__bytes_write.sum"]
        t, tu = d.get("gpu__time_duration.sum", (0.0, "ms"))
        out.append({"kernel": d["kernel"], "dram_read_bytes": rd * scale.get(ru, 1.0),
                    "dram_write_bytes": wr * scale.get(wu, 1.0), "ncu_time_ms": t * scale.get(tu, 1.0)})
    avg = sum(o["dram_read_bytes"] + o["dram_write_bytes"] for o in out) / max(1, len(out))
    import re as _re
    m = _re.search(r"--chunk (\d+)", cmd)
    json.dump({"source": cmd, "chunk": int(m.group(1)) if m else None, "launches": out,
               "avg_dram_bytes_per_launch": avg}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    mode, path, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    {"launches": launches, "traffic": traffic}[mode](path, cmd)
