#!/bin/bash
# first GPU pass of the round: GPU suite, default bench, launch list and full ncu captures of the hot kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02a_launches_csg64_1024.csv python scripts/profile_step.py csg64_1024 > gpurun_out/r02a_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_warp_codes|k_classify_cells|k_vertex_count|k_vertex_emit|k_tet_emit|k_bisect_edges_tiled|k_sdf_field_tiled|k_tile_decide|k_subbox_decide|k_scan_tiles|k_tet_quality_list|k_quality_mark' \
    -o gpurun_out/r02a_hot_csg64_1024 -f python scripts/profile_step.py csg64_1024 > gpurun_out/r02a_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
