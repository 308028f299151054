#!/bin/bash
# bench.py at N ranks (N = visible GPUs by default), device timing only unless E2E=1; prints per-rank diagnostics
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${1:-scale}; N=${2:-$(nvidia-smi -L | wc -l)}
EXTRA="--no-cpu-baseline"; [ "${E2E:-0}" = "1" ] || EXTRA="$EXTRA --no-e2e"
if [ "${PYTEST:-0}" = "1" ]; then
  ( time python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} $EXTRA > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/${TAG}_bench_$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"],2), d["stage_ms"], d["config"]["mesh_checksum"])
for r in d.get("per_rank",[]): print("   ", r)
PY
