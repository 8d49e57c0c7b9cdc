mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
for c in c1 c3; do timeout 600 python bench.py --workload $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_$c.json 2> gpurun_out/f_bench_$c.err; done
timeout 600 python bench.py --workload c4 --spp 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
timeout 600 python bench.py --workload c5 --spp 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/f_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_c2_ncu.log 2>&1
# ncu --set full of the wavefront kernels of one chunk (full frame at 16 spp = one 32 Mi-path chunk: 18 launches)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-count 6 -f -o gpurun_out/f_prof_wf_c2 \
  python tools/profile_render.py --workload c2 --spp 16 --reps 1 --variant 6 > gpurun_out/f_prof_wf_c2_ncu.log 2>&1
