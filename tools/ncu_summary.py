"""Key figures of an `ncu -i X.ncu-rep --page raw --csv` dump, one JSON object per profiled launch.
    python tools/ncu_summary.py raw.csv [algorithmic_bytes_per_launch]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
body = [r for r in rows[rows.index(hdr) + 1:] if len(r) >= len(hdr) and r[0].strip().isdigit()]
PICK = {
    "gpu_time_us": ("gpu__time_duration.sum", 1e-3),
    "sm_clock_ghz": ("sm__cycles_elapsed.avg.per_second", 1e-9),
    "dram_bytes_read": ("dram__bytes_read.sum", 1.0),
    "dram_bytes_write": ("dram__bytes_write.sum", 1.0),
    "dram_throughput_pct": ("dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "lts_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l2_hit_rate_pct": ("lts__t_sector_hit_rate.pct", 1.0),
    "tensor_pipe_active_pct_of_elapsed": ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    "issue_active_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_elapsed", 1.0),
    "registers_per_thread": ("launch__registers_per_thread", 1.0),
    "achieved_occupancy_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
}


def col(sub):
    c = [i for i, h in enumerate(hdr) if h.endswith(sub) or h == sub]
    return c[0] if c else None


units = rows[rows.index(hdr) + 1] if rows.index(hdr) + 1 < len(rows) else []
for r in body:
    out = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
    for key, (sub, scale) in PICK.items():
        i = col(sub)
        if i is None or not r[i]:
            continue
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            continue
        u = units[i] if i < len(units) else ""
        # ncu prints byte counts / times in scaled units: normalise the common ones
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0,
                "second": 1e9, "cycle/nsecond": 1e9, "cycle/usecond": 1e6, "cycle/second": 1.0}.get(u, None)
        if key.startswith("dram_bytes") and mult:
            v *= mult
        elif key == "gpu_time_us" and mult:
            v = v * mult * 1e-3
            out[key] = v
            continue
        elif key == "sm_clock_ghz" and mult:
            v = v * mult * 1e-9
            out[key] = v
            continue
        out[key] = v * scale if key not in ("gpu_time_us", "sm_clock_ghz") or not mult else v
    if "dram_bytes_read" in out and "gpu_time_us" in out:
        out["dram_gbs"] = (out["dram_bytes_read"] + out.get("dram_bytes_write", 0.0)) / (out["gpu_time_us"] * 1e-6) / 1e9
    if len(sys.argv) > 2 and "gpu_time_us" in out:
        out["algorithmic_gbs"] = float(sys.argv[2]) / (out["gpu_time_us"] * 1e-6) / 1e9
    print(json.dumps(out))
