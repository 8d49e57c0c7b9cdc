# ncu --set full of the wavefront kernels (first levels) of one workload: bash scripts/gpu_prof_wf.sh TAG WORKLOAD SPP
TAG=$1; WL=$2; SPP=$3; SKIP=${4:-0}; COUNT=${5:-4}
mkdir -p gpurun_out
timeout 300 python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant 6 > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
# only the kernels without work counters (template argument STATS = false), SKIP launches skipped (12 per chunk at 6 bounces), COUNT captured
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip $SKIP --launch-count $COUNT -f -o gpurun_out/${TAG} \
  python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant 6 > gpurun_out/${TAG}_ncu.log 2>&1
