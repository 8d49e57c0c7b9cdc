mkdir -p gpurun_out
rm -f gpurun_out/t_ab.log
bash scripts/gpu_ab.sh t_ab "permesh merged" "c3 2 3" "c4 4 6" "c5 2 6"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling" > gpurun_out/t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/t_pytest.log
