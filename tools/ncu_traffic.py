"""Distil `ncu --set full --page raw --csv` pages under profiles/ into profiles/ncu_traffic.json.

    python tools/ncu_traffic.py [--dominant profiles/<page>.csv[::kernel-name-substring]]

(a page may hold many launches: the LONGEST launch whose name contains the substring is taken)

For every kernel row of every `profiles/*ncu_full*.csv`: DRAM bytes read / written per launch, duration, achieved DRAM
bandwidth and the tensor-pipe / issue utilisation when the page has them.  bench.py reads the entry marked "dominant" for
`roofline.traffic` (so that number is regenerated from a committed ncu page, never typed in by hand).
"""
from __future__ import annotations

import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
        "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def _num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def parse(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue

        def get(name, scale=True):
            i = col.get(name)
            if i is None:
                return None
            v = _num(r[i])
            if v is None:
                return None
            return v * UNIT.get(units[i], 1.0) if scale else v

        rd, wr, ms = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
        if rd is None and wr is None and ms is None:
            continue
        ent = {"file": os.path.relpath(path, ROOT), "kernel": r[col["Kernel Name"]].replace("void <unnamed>::", "")[:120],
               "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "dram_read_bytes": rd, "dram_write_bytes": wr, "time_ms": ms}
        if rd is not None and wr is not None and ms:
            ent["dram_gbs"] = (rd + wr) / ms / 1e6
        for key, name in [("tensor_pipe_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                          ("tensor_pipe_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
                          ("issue_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                          ("sm_clock_mhz", "sm__cycles_elapsed.avg.per_second"),
                          ("l2_hit_pct", "lts__t_sector_hit_rate.pct")]:
            v = get(name, scale=False)
            if v is not None and key not in ent:
                if key == "sm_clock_mhz" and units[col[name]].lower() == "ghz":
                    v *= 1000.0
                ent[key] = v
        out.append(ent)
    return out


def main():
    dominant, dom_sub = None, ""
    if "--dominant" in sys.argv:
        spec = sys.argv[sys.argv.index("--dominant") + 1]
        spec, _, dom_sub = spec.partition("::")
        dominant = os.path.relpath(os.path.abspath(spec), ROOT)
    dst = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    prev = json.load(open(dst)) if os.path.exists(dst) else {}
    if dominant is None:
        dominant, dom_sub = prev.get("dominant_file"), prev.get("dominant_kernel_substring", "")
    entries = []
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full*.csv"))):
        entries += parse(p)
    cands = [e for e in entries if e["file"] == dominant and dom_sub in e["kernel"] and e.get("time_ms")]
    dom = max(cands, key=lambda e: e["time_ms"]) if cands else None
    json.dump({"how": "tools/ncu_traffic.py over profiles/*ncu_full*.csv (ncu --set full --clock-control none, --page raw --csv)",
               "dominant_file": dominant, "dominant_kernel_substring": dom_sub, "dominant": dom, "kernels": entries}, open(dst, "w"), indent=1)
    print(f"{len(entries)} kernel pages -> {dst}; dominant = {dom['kernel'] if dom else None}")


if __name__ == "__main__":
    main()
