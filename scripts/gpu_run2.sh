set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale" > gpurun_out/r2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest.log
V=""; for t in 1 4 8 12 16 20 24 28 32; do V="$V $((5 + (t<<20)))"; done
timeout 600 python tools/variance_probe.py c2 8 3 $V > gpurun_out/r2_sweep_c2.log 2>&1
timeout 600 python tools/variance_probe.py c5 2 3 5 $((5 + (8<<20))) $((5 + (24<<20))) > gpurun_out/r2_sweep_c5.log 2>&1
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c2.json 2> gpurun_out/r2_c2.err
