# pond scene (config 3) forced to the wavefront: plain times of variants 3 and 6, then per-kernel durations of variant 6
mkdir -p gpurun_out
python tools/profile_render.py --workload c3 --spp 2 --reps 2 --variant 3 > gpurun_out/h_c3_v3.log 2>&1
python tools/profile_render.py --workload c3 --spp 2 --reps 2 --variant 6 > gpurun_out/h_c3_v6.log 2>&1
bash scripts/gpu_times_wf.sh h_c3_v6_kernels c3 2 6
