"""Key figures of an `ncu -i X.ncu-rep --page raw --csv` dump, one JSON object per profiled launch.
    python tools/ncu_summary.py raw.csv [algorithmic_bytes_per_launch]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
body = [r for r in rows[hi + 2:] if len(r) >= len(hdr) and r[0].strip().isdigit()]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0,
         "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "Ghz": 1.0, "Mhz": 1e-3, "%": 1.0, "": 1.0,
         "sector": 32.0, "register/thread": 1.0, "cycle": 1.0}
PICK = {   # key: (exact metric name, what the scaled value is)
    "gpu_time_us": "gpu__time_duration.sum",
    "sm_clock_ghz": "sm__cycles_elapsed.avg.per_second",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "l2_to_sm_read_bytes": "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "tensor_pipe_active_pct_of_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "registers_per_thread": "launch__registers_per_thread",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
for r in body:
    out = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
    for key, name in PICK.items():
        if name not in hdr:
            continue
        i = hdr.index(name)
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            continue
        out[key] = v * SCALE.get(units[i], 1.0)
    if "dram_bytes_read" in out and "gpu_time_us" in out:
        out["dram_gbs"] = (out["dram_bytes_read"] + out.get("dram_bytes_write", 0.0)) / out["gpu_time_us"] / 1e3
    if len(sys.argv) > 2 and "gpu_time_us" in out:
        out["algorithmic_gbs"] = float(sys.argv[2]) / out["gpu_time_us"] / 1e3
    print(json.dumps(out))
