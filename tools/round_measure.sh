#!/bin/bash
# End-of-round measurement set (GPU box): tests, default bench, reference arm, nearest-mode benches, ncu launch list.
# Usage: gpurun --timeout 1500 -- 'bash tools/round_measure.sh r1_r'
tag=${1:-rX}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.log 2> $out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_ref.log 2>&1
python bench.py --workload b64_2048 --interp nearest --steps 10 --no-e2e --no-cpu-baseline > $out/bench_${tag}_nearest2048.log 2>&1
python bench.py --interp nearest --steps 20 --no-e2e --no-cpu-baseline > $out/bench_${tag}_nearest1024.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_$tag.log 2>&1
python - <<PY
import json
for f in ("bench_$tag", "bench_${tag}_nearest2048", "bench_${tag}_nearest1024"):
    d = json.loads(open("$out/" + f + ".log").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"], 3), round(d["serial_ms_per_step"], 3), round(d["roofline"]["frac"], 4),
          d.get("e2e", {}).get("value"), round(d["mask_mode"]["serial_ms_per_step"], 3), d["mask_mode"].get("c1_tail"))
PY
tail -c 300 $out/bench_${tag}_ref.log
