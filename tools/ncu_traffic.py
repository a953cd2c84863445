"""Record the DRAM traffic `ncu --set full` measured for one kernel launch in profiles/ncu_traffic.json, the file
bench.py reads `roofline.traffic` from (kernels without an entry report null).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep --launch 0 --alg-bytes 301989888 \
        --note "24x24x64 -> 24x24x64 block, batch 1024" --details profiles/r2_ncu_xxx_details.txt

The kernel name is normalised to the profiler's naming (`tcp_dwpw_kernel<3,1>`).  Runs in the build container (ncu reads
reports without a GPU)."""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--launch", type=int, default=0, help="index of the launch inside the report")
ap.add_argument("--alg-bytes", type=float, required=True, help="algorithmic bytes of THAT launch (DESIGN.md traffic model)")
ap.add_argument("--note", default="")
ap.add_argument("--details", default="", help="also write `ncu --page details` of the report to this file")
args = ap.parse_args()

raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
body = [r for r in rows[2:] if len(r) == len(H)]
r = body[args.launch]
col = {h: i for i, h in enumerate(H)}


def val(name):
    return float(r[col[name]].replace(",", ""))


name = r[col["Kernel Name"]]
short = re.sub(r"\(int\)|\(bool\)|\s", "", name)
short = re.sub(r"^.*?(\w+_kernel)(<[^>]*>)?.*$", lambda m: m.group(1) + (m.group(2) or ""), short)
units = {h: rows[1][i] for i, h in enumerate(H)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = val("dram__bytes_read.sum") * scale[units["dram__bytes_read.sum"]]
wr = val("dram__bytes_write.sum") * scale[units["dram__bytes_write.sum"]]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
db[short] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": args.alg_bytes,
             "ratio": (rd + wr) / args.alg_bytes, "source": args.details or os.path.basename(args.report), "note": args.note,
             "duration_us_under_ncu": val("gpu__time_duration.sum") / (1e3 if units["gpu__time_duration.sum"] in ("nsecond", "ns") else 1.0)}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(short, json.dumps(db[short]))
if args.details:
    det = subprocess.run(["ncu", "-i", args.report, "--page", "details"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(ROOT, args.details), "w").write(det)
