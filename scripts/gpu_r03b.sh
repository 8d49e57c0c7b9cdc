# round 2, call 3b (1 GPU): parity of the slab-test trace build, A/B against the cone-test build, then source-level ncu hot
# spots of the pond scene (config 3: state-machine kernel, and the wavefront forced on it) and of config 5's sample kernel.
# The reports are summarised ON the box (scripts/ncu_hotspots.py) and deleted: gpurun_out/ may not exceed 64 MiB.
mkdir -p gpurun_out/r03b
O=gpurun_out/r03b
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
bash scripts/gpu_ab.sh r03b/ab_lc_slab "lc_cone" "c2 16 0" "c4 4 0" "c5 2 0"
echo "=== default (slab)" >> $O/ab_lc_slab.log
for args in "c2 16 0" "c4 4 0" "c5 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> $O/ab_lc_slab.log; done
cat $O/ab_lc_slab.log
LIB=hai719-raytracing_b200/lib/libhai719_rt.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_regen --launch-skip 0 --launch-count 1 -f -o /tmp/c3_regen \
  python tools/profile_render.py --workload c3 --spp 1 --reps 1 --no-stats > $O/c3_regen_ncu.log 2>&1
python scripts/ncu_hotspots.py /tmp/c3_regen.ncu-rep $LIB --kernel k_render_regen --top 70 > $O/c3_regen_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c3_regen.ncu-rep $LIB --kernel k_render_regen --top 40 --depth 2 > $O/c3_regen_hotspots_d2.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 0 --launch-count 7 -f -o /tmp/c3_wf \
  python tools/profile_render.py --workload c3 --spp 1 --reps 1 --no-stats --variant 6 > $O/c3_wf_ncu.log 2>&1
ncu -i /tmp/c3_wf.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active > $O/c3_wf_raw.csv 2>&1
for i in 0 1; do python scripts/ncu_hotspots.py /tmp/c3_wf.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)2>" --index $i --top 50 > $O/c3_wf_light2_${i}_hotspots.txt 2>&1; done
python scripts/ncu_hotspots.py /tmp/c3_wf.ncu-rep $LIB --kernel "k_wf_trace" --index 0 --top 50 > $O/c3_wf_trace_0_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c3_wf.ncu-rep $LIB --kernel "k_wf_trace" --index 1 --top 50 > $O/c3_wf_trace_1_hotspots.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 0 --launch-count 4 -f -o /tmp/c5_wf \
  python tools/profile_render.py --workload c5 --spp 1 --reps 1 --no-stats > $O/c5_wf_ncu.log 2>&1
ncu -i /tmp/c5_wf.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active > $O/c5_wf_raw.csv 2>&1
for i in 0 1; do python scripts/ncu_hotspots.py /tmp/c5_wf.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)2>" --index $i --top 50 > $O/c5_wf_light2_${i}_hotspots.txt 2>&1; done
cp /tmp/c3_regen.ncu-rep $O/
ls -la $O
