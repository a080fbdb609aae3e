#!/bin/bash
# A/B of FOVEA_FILL_PAD_KB (a cap on the fill kernel's CTAs per SM through unused dynamic shared memory) under the pipelined
# schedule: value, pipelined / serial ms per step, the fill alone and overlapped.
for kb in ${@:-0 56 75 110}; do
  FOVEA_FILL_PAD_KB=$kb python bench.py --steps 30 --warmup 3 --no-e2e --no-extras --no-cpu-baseline 2>/dev/null > /tmp/pad_$kb.json
  python - "$kb" <<'PY'
import json, sys
kb = sys.argv[1]
d = json.loads(open(f"/tmp/pad_{kb}.json").read().strip().split("\n")[-1])
r = d["roofline"]
print("pad_kb", kb, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "serial", round(d["serial_ms_per_step"], 3),
      "fill alone", round(r["ms_per_launch"], 3), "overlapped", round(r["ms_per_launch_overlapped"], 3))
PY
done
