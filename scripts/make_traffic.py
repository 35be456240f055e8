"""profiles/traffic.json from ncu --set full reports: per profiled kernel name (the names bench.py's per-kernel timing
uses) the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the L2 bytes (lts__t_sectors.sum x 32 B) of ONE
launch.  usage: python scripts/make_traffic.py out.json report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

# (one build launches the envelope kernel, the resolve pass and the fused fill twice each -- scene bands / rows on the second
# stream first, then the far ones: the n-th launch of a report maps to the n-th name)
NAMES = {"search_warp_kernel": "search", "dt_row_band_kernel": ["dt_row_envelope", "dt_row_envelope_far"],
         "dt_fill_propagate_kernel": ["dt_fill_propagate", "dt_fill_propagate_far"], "integral_tma_kernel": "integral",
         "dt_resolve_kernel": ["dt_resolve", "dt_resolve_far"], "dt_col_band_kernel": "dt_col_band", "dt_l1_propagate_kernel": "dt_l1_propagate",
         "raster_kernel": "raster", "topk_level1_kernel": "topk", "search_key_kernel": "search_order"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1.0}
out, src = {}, []
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "gpu__time_duration.sum")}
    for r in rows[2:]:
        kname = r[col["Kernel Name"]]
        key = next((v for k, v in NAMES.items() if k in kname), None)
        if isinstance(key, list):
            key = next((k for k in key if k not in out), None)
        if key is None or key in out:
            continue
        val = lambda c: float(r[col[c]].replace(",", "")) * UNIT.get(units[col[c]], 1.0)
        out[key] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "l2_bytes": val("lts__t_sectors.sum") * 32.0,
                    "ncu_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[col["gpu__time_duration.sum"]], 1.0),
                    "report": rep.split("/")[-1]}
    src.append(rep.split("/")[-1])
out["_source"] = "ncu --set full --clock-control none, one launch per kernel: dram__bytes_read.sum + dram__bytes_write.sum, lts__t_sectors.sum x 32 B; reports: " + ", ".join(src)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
