mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
bash scripts/gpu_times_wf.sh k_times_c2 c2 16
