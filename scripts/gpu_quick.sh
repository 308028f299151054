#!/bin/bash
# quick GPU check: parity suite (+ poison mode on the small cases) and a short bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-quick}
( time python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
( MC_DEBUG=2 python -m pytest tests/test_gpu_parity.py tests/test_slabs.py -m gpu -q -x -k "not full_size" ) > gpurun_out/${TAG}_pytest_poison.log 2>&1
echo "poison rc=$?" >> gpurun_out/${TAG}_pytest_poison.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_pytest_poison.log
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], d["config"]["mesh_checksum"])
PY
