# issue-side utilisation of EVERY render kernel of one frame: bash scripts/gpu_issue.sh TAG WORKLOAD SPP [VARIANT]
# one ncu pass set: duration, issue-slot utilisation, active lanes per warp instruction, DRAM bytes. Feeds profiles/latest.json
# (scripts/ncu_issue_summary.py). A number printed by the profiled program is never a bench value.
TAG=$1; WL=$2; SPP=$3; VAR=${4:-0}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --kernel-name-base mangled -k regex:"k_wf_|k_camera|k_pixel|k_resolve|k_render_regen|k_render_paths" -c 400 --csv --log-file gpurun_out/${TAG}.csv \
  python tools/profile_render.py --workload $WL --spp $SPP --reps 1 --variant $VAR --no-stats > gpurun_out/${TAG}.log 2>&1
