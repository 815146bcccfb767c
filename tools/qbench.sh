#!/bin/bash
# quick developer bench: no CPU sample, prints the few numbers that matter
timeout 200 python bench.py --steps ${1:-10} --cpu-batch 0 > gpurun_out/qb.json 2> gpurun_out/qb.err; echo bench rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/qb.json"))
print("value %.2fM  ms/step %.2f  K1 %.4f ms frac %.3f  e2e %.2fM" % (d["value"] / 1e6, d["ms_per_step"],
      d["roofline"]["launch_ms"], d["roofline"]["frac"], d["e2e"]["value"] / 1e6))
print(d["breakdown_ms_per_step"])
PY
