mkdir -p gpurun_out
V=""; for t in 1 8 16 24 32; do V="$V $((5 + (t<<20)))"; done
timeout 600 python tools/variance_probe.py c2 8 3 5 $((5 + (1<<28))) $V 2>&1 | grep -v "^upload [12]" > gpurun_out/r4_sweep_c2.log
timeout 600 python tools/variance_probe.py c5 2 3 5 $((5 + (1<<28))) 2>&1 | grep -v "^upload [12]" > gpurun_out/r4_sweep_c5.log
timeout 600 python tools/variance_probe.py c3 2 0 $((1<<28)) 2>&1 | grep -v "^upload [12]" > gpurun_out/r4_sweep_c3.log
timeout 600 python tools/variance_probe.py c4 2 0 $((1<<28)) 2>&1 | grep -v "^upload [12]" > gpurun_out/r4_sweep_c4.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4_pytest.log
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r4_c2.json 2> gpurun_out/r4_c2.err
