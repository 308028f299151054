#!/bin/bash
# BASELINE configs 2 and 3 (and the 512^3 / 256-primitive variants) as bench lines on one GPU
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; TAG=${1:-wl}
: > gpurun_out/${TAG}_workloads.jsonl
for WL in sphere_256 xplicit_512 csg64_512 csg256_1024; do
  python bench.py --workload $WL --steps 5 --warmup 3 >> gpurun_out/${TAG}_workloads.jsonl 2> gpurun_out/${TAG}_${WL}.err
  echo "$WL rc=$?"; tail -2 gpurun_out/${TAG}_${WL}.err
done
python - <<PY
import json
for l in open("gpurun_out/${TAG}_workloads.jsonl"):
    d=json.loads(l); print(d["config"]["workload"][:60], round(d["ms_per_step"],2), d["stage_ms"], d["config"]["output_tets"], d["config"]["boundary_sides"], d["config"]["side_sets"], d.get("cpu_baseline",{}).get("value"))
PY
