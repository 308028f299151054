#!/bin/bash
# GPU pass: GPU suite, poison run, default bench, launch list and full ncu captures of the hot kernels
# (reports are exported to CSV on the box: gpurun_out/ is limited to 64 MiB)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
free -g >> gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
( MC_DEBUG=2 python -m pytest tests/test_gpu_parity.py tests/test_slabs.py -m gpu -q -x -k "not full_size" ) > gpurun_out/${TAG}_pytest_poison.log 2>&1
echo "poison rc=$?" >> gpurun_out/${TAG}_pytest_poison.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches_csg64_1024.csv python scripts/profile_step.py csg64_1024 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_warp_codes|k_classify_cells|k_vertex_count|k_vertex_emit|k_tet_emit|k_bisect_edges_tiled|k_sdf_field_tiled|k_tile_decide|k_subbox_decide|k_tet_quality_list|k_quality_mark|k_boundary_sides|k_sideset|k_eval_side|k_mark_side' \
    -o /tmp/${TAG}_hot -f python scripts/profile_step.py csg64_1024 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > gpurun_out/${TAG}_hot_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_hot.ncu-rep --page details --csv > gpurun_out/${TAG}_hot_details.csv 2>/dev/null
for k in k_warp_codes k_classify_cells k_vertex_count k_vertex_emit k_tet_emit_full k_tet_emit k_sdf_field_tiled k_boundary_sides; do
  ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv --kernel-name regex:"^$k$" > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
done
xz -T4 -f gpurun_out/${TAG}_hot_raw.csv gpurun_out/${TAG}_hot_details.csv gpurun_out/${TAG}_src_*.csv
du -sh gpurun_out; ls -la gpurun_out | tail -30
