#!/bin/bash
# multi-GPU pass: real-rank parity tests, then bench.py under torchrun at every N <= visible GPUs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-multi}
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
( time python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_bench_1.json 2> gpurun_out/${TAG}_bench_1.err
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
    echo "bench N=$N rc=$?"
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],2), d["config"]["mesh_checksum"], d["e2e"]["ms_per_step"] if d.get("e2e") else None)
        for r in d.get("per_rank",[]): print("   ", r)
    except Exception as e: print(f, "ERR", e)
PY
