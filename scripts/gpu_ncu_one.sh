#!/bin/bash
# ncu --set full capture of selected kernels of one workload; exports raw + source CSV (SASS and CUDA-C views)
# usage: gpu_ncu_one.sh TAG WORKLOAD 'kernel-regex' [kernel names for the source page...]
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=$1; WL=$2; RE=$3; shift 3
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$RE" \
    -o /tmp/${TAG} -f python scripts/profile_step.py $WL > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
for k in "$@"; do
  ncu -i /tmp/${TAG}.ncu-rep --page source --csv --kernel-name regex:"^$k" > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
  ncu -i /tmp/${TAG}.ncu-rep --page source --print-source cuda --csv --kernel-name regex:"^$k" > gpurun_out/${TAG}_cuda_$k.csv 2>/dev/null
done
xz -T4 -f gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_src_*.csv gpurun_out/${TAG}_cuda_*.csv
ls -la gpurun_out | grep ${TAG}
