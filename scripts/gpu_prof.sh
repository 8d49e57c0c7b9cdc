# ncu --set full capture of the render kernel of one workload: plain run first, then the same command under ncu
# usage: bash scripts/gpu_prof.sh TAG WORKLOAD SPP [VARIANT]
TAG=$1; WL=$2; SPP=$3; VAR=${4:-0}
mkdir -p gpurun_out
timeout 300 python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant $VAR > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 --launch-count 1 -f -o gpurun_out/${TAG} \
  python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant $VAR > gpurun_out/${TAG}_ncu.log 2>&1
