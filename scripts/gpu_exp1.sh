#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for MF in 256 16; do for WL in sphere_256 xplicit_512; do
  MC_TILED_MIN_FLOPS=$MF python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$MF','$WL',round(d['ms_per_step'],2),d['stage_ms'],d['config']['mesh_checksum']['tets'])"
done; done
