#!/bin/bash
# End-of-round measurement set (GPU box, one GPU): tests, default bench, reference arm, 2048^2 / nearest benches, the ncu
# launch list of one bench command and `ncu --set full` captures of the kernels DESIGN.md discusses.
# Usage: gpurun --timeout 2400 -- 'bash tools/round_measure.sh r2_p'
tag=${1:-rX}
out=gpurun_out
python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_reference_arm.json 2>&1
python bench.py --interp nearest --steps 20 --no-e2e --no-cpu-baseline > $out/bench_${tag}_nearest_1024.json 2>&1
python bench.py --triangulation host --steps 3 --no-e2e --no-extras --no-cpu-baseline > $out/bench_${tag}_host_tri.json 2>&1
python tools/bench_stock_cuda.py > $out/${tag}_stock_cuda.json 2>&1
python tools/probe_grid_sample.py b64_1024 > $out/${tag}_probe_grid_sample.txt 2>&1
python tools/probe_grid_sample.py b64_2048 >> $out/${tag}_probe_grid_sample.txt 2>&1
python tools/probe_delaunay_frames.py b64_1024 > $out/${tag}_delaunay_frames.txt 2>&1
python tools/probe_mask.py > $out/${tag}_probe_mask.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $out/ncu_$tag.log 2>&1
for k in inverse_fill_kernel raster_mark_kernel raster_fill_rows_kernel delaunay_kernel inverse_mask_kernel triangle_candidates_kernel inverse_fill_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $out/${tag}_$k \
      python tools/probe_kernels.py > $out/ncu_${tag}_$k.log 2>&1
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_ncu_raw.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:grid_sample_fwd_kernel -s 2 -c 1 -o $out/${tag}_grid_sample_fwd python tools/probe_grid_sample.py b64_1024 > /dev/null 2>&1
ncu -i $out/${tag}_grid_sample_fwd.ncu-rep --page raw --csv > $out/${tag}_grid_sample_fwd_ncu_raw.csv 2>/dev/null
FOVEA_GS_TMA=1 ncu --set full --clock-control none -k regex:grid_sample_fwd_tma_kernel -s 2 -c 1 -o $out/${tag}_grid_sample_fwd_tma python tools/probe_grid_sample.py b64_1024 > /dev/null 2>&1
ncu -i $out/${tag}_grid_sample_fwd_tma.ncu-rep --page raw --csv > $out/${tag}_grid_sample_fwd_tma_ncu_raw.csv 2>/dev/null
rm -f $out/${tag}_*.ncu-rep
python - <<PY
import json
for f in ("bench_$tag", "bench_${tag}_nearest_1024", "bench_${tag}_host_tri"):
    d = json.loads(open("$out/" + f + ".json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"], 3), round(d["serial_ms_per_step"], 3), round(d["roofline"]["frac"], 4),
          d.get("e2e", {}).get("value"), round(d["mask_mode"]["serial_ms_per_step"], 3))
PY
tail -c 300 $out/bench_${tag}_reference_arm.json
