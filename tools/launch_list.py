"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` log of tools/profile_step.py to the launches of the LAST
pipeline pass (from the last 5x5 stem kernel on) and print per-kernel-class totals.

    python tools/launch_list.py gpurun_out/launches_all.csv profiles/rX_launches_batch1024.csv
"""
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
body = [r for r in rows[hdr + 2:] if len(r) > vi]
starts = [i for i, r in enumerate(body) if re.search(r"stem(_mma)?_kernel<(\(int\))?5", r[ki])]
last = body[starts[-1]:]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(H)
    w.writerows(last)
tot = 0.0
classes = {}
for r in last:
    us = float(r[vi].replace(",", "")) / 1000.0
    name = re.sub(r"^.*?(\w+_kernel|\w+)<.*$", r"\1", r[ki]) if "<" in r[ki] else r[ki].split("(")[0].split("::")[-1]
    classes[name] = classes.get(name, 0.0) + us
    tot += us
print(f"{len(last)} launches in the last pass, {tot / 1000.0:.3f} ms under ncu (cold caches, serialised)")
for k, v in sorted(classes.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v / 1000.0:7.3f} ms  {100.0 * v / tot:5.1f} %")
