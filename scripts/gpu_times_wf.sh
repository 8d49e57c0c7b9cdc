# per-kernel durations (ncu, gpu__time_duration only) of one render: bash scripts/gpu_times_wf.sh TAG WORKLOAD SPP [VARIANT]
TAG=$1; WL=$2; SPP=$3; VAR=${4:-6}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:"k_wf_.*ILb0|k_camera|k_resolve|k_render_regenILb0" -c 200 --csv --log-file gpurun_out/${TAG}.csv \
  python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant $VAR > gpurun_out/${TAG}.log 2>&1
