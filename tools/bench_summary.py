"""Print the interesting parts of bench.py JSON lines: python tools/bench_summary.py gpurun_out/bench4.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    r = d.get("roofline", {})
    print(f"{f}: {d['value']:.0f} {d['unit']}  ms/step {d.get('ms_per_step', 0):.3f}  e2e {d['e2e']['value']:.0f}  launches {d.get('gpu_launches')}")
    if "all_frames_landmarked" in d:
        print(f"   all_frames_landmarked {d['all_frames_landmarked']['value']:.0f} ({d['all_frames_landmarked']['ms_per_step']:.3f} ms)")
    print(f"   roofline: {r.get('kernel')} x{r.get('launches')} frac {r.get('frac')} share {r.get('share_of_step')}  pipeline {r.get('pipeline', {}).get('frac')}")
    for k, v in d.get("kernels", {}).items():
        print(f"     {k:24s} n={v['launches']:3d} ms={v['ms']:.4f} share={v['share']:.3f} GB/s={v['GBps']} TF={v['TFLOPs']}")
        for fn, fv in v.get("functions", {}).items():
            print(f"         {fn:44s} n={fv['launches']:3d} ms={fv['ms']:.4f} GB/s={fv['GBps']}")
