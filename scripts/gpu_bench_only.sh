#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; TAG=${1:-b}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], d["config"]["mesh_checksum"], d["config"]["boundary_sides"], d["config"]["side_sets"])
print(d["e2e"])
PY
