#!/bin/bash
# Development tool: summaries of the ncu reports that tools/profile_r02.sh left in gpurun_out/ -> profiles/ (tracked).
set -u
cd "$(dirname "$0")/.."
for f in gpurun_out/prof_*_r02.ncu-rep; do
    n=$(basename "$f" .ncu-rep)
    { python tools/ncu_summary.py "$f"; echo; echo "== hottest SASS lines (warp-stall samples)"; python tools/ncu_source_hot.py "$f" 25 | cut -c1-170; } > "profiles/${n}_summary.txt" 2>/dev/null
done
python tools/launch_list_summary.py gpurun_out/r02_launches.csv > profiles/launches_r02_64replicas_one_lane.txt
python tools/launch_list_summary.py gpurun_out/r02_launches_r1.csv > profiles/launches_r02_1replica.txt
python - <<'PY'
import csv, io, json, subprocess
fam = {"update_slice": "prof_window_r02", "update_build_xy": "prof_gather_r02", "update_flush": "prof_flush_r02",
       "cb_mult": "prof_cb_r02", "gemm_dmma": "prof_gemm_r02", "qrcp_factor": "prof_panel_r02"}
out = {}
for name, rep in fam.items():
    try:
        txt = subprocess.run(["ncu", "-i", "gpurun_out/%s.ncu-rep" % rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        h, u, r = rows[0], rows[1], rows[2]
        def val(k):
            v = float(r[h.index(k)].replace(",", ""))
            unit = u[h.index(k)].lower()
            return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        out[name] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    except Exception as e:
        print("skip", name, e)
out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture per family, 64 replicas in one "
                "lane (qrcp_factor: the panel kernel only)")
json.dump(out, open("profiles/traffic_r02.json", "w"), indent=1)
print(out)
PY
