"""profiles/r2_step_traffic.json from the raw CSV page of the one-step ncu capture:
  ncu -i gpurun_out/r2_step.ncu-rep --page raw --csv > gpurun_out/r2_step_raw.csv
  python tools/traffic_from_ncu.py gpurun_out/r2_step_raw.csv 256 profiles/r2_step_traffic.json
Stamped with bench.source_hash(): bench.py reports `roofline.traffic` from it only while the kernel sources are unchanged."""
import sys, os, csv, json, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def num(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else 0.0


def main():
    src, agents, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
             "nsecond": 1e-3, "msecond": 1e3}
    def val(r, key):
        i = col[key]
        return num(r[i]) * scale.get(units[i], 1.0)
    per, shares = [], {}
    rd = wr = tot = 0.0
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("saceo::", "").replace("void ", "")
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        us = val(r, "gpu__time_duration.sum")
        b_r, b_w = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        per.append([name, grid, round(us, 2), int(b_r), int(b_w)])
        shares[name] = shares.get(name, 0.0) + us
        rd += b_r; wr += b_w; tot += us
    out = {"agents": agents, "source_hash": bench.source_hash(), "launches_per_step": len(per),
           "serialized_us": round(tot, 3), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
           "note": "ncu --set full --clock-control none, one un-graphed update (tools/step_traffic.py); per-launch times are "
                   "serialised and cold-cache: only the shares are comparable with bench.py's live per-kernel times",
           "shares_us": [[k, round(v, 3)] for k, v in shares.items()],
           "per_launch": per}
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, "launches", len(per), "us", round(tot, 1), "GB", round((rd + wr) / 1e9, 3))


if __name__ == "__main__":
    main()
