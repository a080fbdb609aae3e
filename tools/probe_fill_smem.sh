#!/bin/bash
# A/B of FOVEA_FILL_SMEM (table rows of a tile staged in shared memory) -- the fill alone and under the pipelined schedule.
for v in 0 1 2; do
  FOVEA_FILL_SMEM=$v python bench.py --steps 30 --warmup 3 --no-e2e --no-extras --no-cpu-baseline 2>/dev/null > /tmp/smem_$v.json
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open(f"/tmp/smem_{v}.json").read().strip().split("\n")[-1])
r = d["roofline"]
print("FOVEA_FILL_SMEM", v, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "serial", round(d["serial_ms_per_step"], 3),
      "fill alone", round(r["ms_per_launch"], 3), "overlapped", round(r["ms_per_launch_overlapped"], 3), "frac", round(r["frac"], 4))
PY
done
