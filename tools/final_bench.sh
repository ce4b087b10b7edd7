#!/bin/bash
# final numbers of a round: headline with default flags (incl. cpu_baseline), the other configs without the CPU leg
python bench.py > gpurun_out/final_cornell_monkey.json 2> gpurun_out/final_cornell_monkey.err
for sc in cornell_boxes matball mega; do
  python bench.py --scene $sc --steps 10 --warmup 3 --no-cpu > gpurun_out/final_$sc.json 2> gpurun_out/final_$sc.err
done
for sc in cornell_monkey cornell_boxes matball mega; do
python - <<PY
import json
d=json.loads(open("gpurun_out/final_$sc.json").read().strip().splitlines()[-1])
print("$sc", "value %.1f  ms %.3f  e2e %.1f  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("   stage", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "roofline frac %.3f l2 %.3f" % (d["roofline"]["frac"], d["roofline"]["frac_of_l2"]))
PY
done
