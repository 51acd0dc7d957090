"""profiles/r2_traffic.json from the ncu --set full raw pages of this round (profiles/r2_<kernel>_raw.csv): DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernels, which bench.py reports as
roofline.traffic.  coarse_solve = two chain launches (forward + backward triangle)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

def rows_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))

def dram(row, units):
    return sum(float(row[k]) * UNIT[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))

out = {"workload": "gl32", "source": "ncu --set full --clock-control none, tools/gpu_r2e.sh"}
r, u = rows_of(os.path.join(ROOT, "profiles", "r2_gs_cluster_raw.csv"))
out["gs_fine"] = {"bytes": dram(r[0], u), "kernel": "k_gs_cluster", "duration_us": float(r[0]["gpu__time_duration.sum"]) * (1e3 if u["gpu__time_duration.sum"] == "ms" else 1)}
r, u = rows_of(os.path.join(ROOT, "profiles", "r2_band_chain_raw.csv"))
out["coarse_solve"] = {"bytes": 2 * dram(r[0], u), "kernel": "2 x k_band_chain"}
r, u = rows_of(os.path.join(ROOT, "profiles", "r2_apply_raw.csv"))
fine = max(r, key=lambda q: float(q["launch__grid_size"]))
out["apply_fine"] = {"bytes": dram(fine, u), "kernel": "k_apply (fine level)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(out)
