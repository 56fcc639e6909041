"""DRAM bytes per launch of the fused block kernel, from `ncu --set full` captures of the benchmark command
(scripts/profile_forward.sh): reads profiles/<tag>_ncu_full_*_resblock2_{precise,fast}.csv (the `--page raw --csv`
export) and writes profiles/r2_resblock_traffic.json, which bench.py reports as `roofline.traffic` -- measured by ncu on
this kernel at this workload, never a constant in bench.py.

    python scripts/ncu_traffic.py profiles/r2_ncu_full_v1_resblock2_precise.csv profiles/r2_ncu_full_v1_resblock2_fast.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    by = lambda k: float(d[k][0]) * UNIT[d[k][1]]
    return {"kernel": d["Kernel Name"][0],
            "dram_bytes_per_launch": by("dram__bytes_read.sum") + by("dram__bytes_write.sum"),
            "dram_read_bytes": by("dram__bytes_read.sum"), "dram_write_bytes": by("dram__bytes_write.sum"),
            "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"][0]),
            "tensor_pipe_pct_active": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
            "tensor_pipe_pct_elapsed": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
            "registers_per_thread": int(float(d["launch__registers_per_thread"][0])),
            "source": os.path.relpath(path, ROOT)}


if __name__ == "__main__":
    out = {}
    for path in sys.argv[1:]:
        e = read(path)
        out["precise" if "ELb1EEE" in e["kernel"] or "(bool)1>" in e["kernel"] or "precise" in path else "fast"] = e
    dest = os.path.join(ROOT, "profiles", "r2_resblock_traffic.json")
    json.dump(out, open(dest, "w"), indent=1)
    print(json.dumps(out, indent=1))
