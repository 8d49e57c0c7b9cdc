mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale" > gpurun_out/r7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r7_pytest.log
timeout 600 python tools/variance_probe.py c2 8 5 6 2>&1 | grep -v "^upload [12]" > gpurun_out/r7_sweep.log
timeout 600 python tools/variance_probe.py c5 2 5 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/r7_sweep.log
timeout 600 python tools/variance_probe.py c3 2 0 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/r7_sweep.log
timeout 600 python tools/variance_probe.py c4 2 0 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/r7_sweep.log
timeout 600 python tools/variance_probe.py c1 1 0 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/r7_sweep.log
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --variant 6 > gpurun_out/r7_c2_v6.json 2> gpurun_out/r7_c2_v6.err
