mkdir -p gpurun_out
rm -f gpurun_out/p_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale or chunk" > gpurun_out/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p_pytest.log
bash scripts/gpu_ab.sh p_ab "lc16 lc32" "c3 2 6 3" "c5 2 6" "c2 32 6"
