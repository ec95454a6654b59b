#!/bin/bash
# region-kernel (frame_boxes) duration under ncu + chained step time, both workloads
for w in c2 c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:frame_boxes -c 12 --csv --log-file gpurun_out/fb_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --cpu-steps 1 > /dev/null 2>&1
  python - $w <<'PY'
import csv, sys
w = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/fb_{w}.csv")) if len(r) > 10 and r[0].isdigit() and "<0, 1>" in r[4]]
t = [float(r[-1]) / 1e3 for r in rows]
print(w, "frame_boxes<0,1> under ncu: n", len(t), "median %.2f us" % sorted(t)[len(t) // 2])
PY
done
python tools/pdl_probe2.py | grep "disable_pdl=0 cycle=True"
