mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale or output_stage" > gpurun_out/r14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r14_pytest.log
rm -f gpurun_out/r14_sweep.log
for wl in "c2 8 6" "c5 2 6" "c4 2 6" "c3 2 6 3" "c1 1 6 1"; do
timeout 600 python tools/variance_probe.py $wl 2>&1 | grep -v "^upload [12]" >> gpurun_out/r14_sweep.log
done
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r14_c2.json 2> gpurun_out/r14_c2.err
