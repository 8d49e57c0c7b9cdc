# round 2, call 4f (1 GPU): source-level hot spots of the pool scene's analytic trace kernel at bounce 1 with the flat box test
mkdir -p gpurun_out/r04f
O=gpurun_out/r04f
LIB=hai719-raytracing_b200/lib/libhai719_rt.so
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 3 --launch-count 1 -f -o /tmp/c4_l1 \
  python tools/profile_render.py --workload c4 --spp 1 --reps 1 --no-stats > $O/c4_l1_ncu.log 2>&1
python scripts/ncu_hotspots.py /tmp/c4_l1.ncu-rep $LIB --kernel "k_wf_trace" --index 0 --top 60 > $O/c4_l1_trace_a_flat_hotspots.txt 2>&1
head -50 $O/c4_l1_trace_a_flat_hotspots.txt
