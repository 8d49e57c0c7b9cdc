# DRAM bytes and duration of every kernel of one chunk of config 2 (full frame at 16 spp = one 33.2 M-path chunk)
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base mangled \
  -k regex:"k_wf_.*ILb0|k_camera|k_pixel|k_resolve" -c 100 --csv --log-file gpurun_out/w_traffic.csv \
  python tools/profile_render.py --workload c2 --spp 16 --reps 1 --variant 6 > gpurun_out/w_traffic.log 2>&1
