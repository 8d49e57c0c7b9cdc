mkdir -p gpurun_out
rm -f gpurun_out/r9_sweep.log
for wl in "c2 8" "c5 2" "c3 2" "c4 2" "c1 1"; do
timeout 600 python tools/variance_probe.py $wl 0 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/r9_sweep.log
done
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --variant 6 > gpurun_out/r9_c2_v6.json 2> gpurun_out/r9_c2_v6.err
bash scripts/gpu_prof_wf.sh r9_prof_wf_c2 c2 8
