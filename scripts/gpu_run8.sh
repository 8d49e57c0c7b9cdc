mkdir -p gpurun_out
bash scripts/gpu_ab.sh r13_ab "x_all y_noww" "c2 8 6 5" "c5 2 6" "c4 2 6" "c3 2 6 3" "c1 1 6 1"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale or output_stage" > gpurun_out/r13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r13_pytest.log
