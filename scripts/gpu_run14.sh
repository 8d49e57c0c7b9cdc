mkdir -p gpurun_out
rm -f gpurun_out/o_sweep.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale or chunk or incremental" > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
U=$((6 + (1<<29)))
for wl in "c2 32 6 $U" "c5 2 6 $U" "c4 4 6 $U"; do timeout 600 python tools/variance_probe.py $wl 2>&1 | grep -v "^upload [12]" >> gpurun_out/o_sweep.log; done
