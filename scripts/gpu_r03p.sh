# round 2, call 3p (1 GPU): box-test tail as one comparison of two min/max: parity + timings; CTAs per SM of the wavefront kernels (6 / 8 / 10) A/B
mkdir -p gpurun_out/r03p
O=gpurun_out/r03p
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
echo "=== default (8 CTAs, new box-test tail)" >> $O/timings.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
bash scripts/gpu_ab.sh r03p/timings "minb6 minb10" "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0"
cat $O/timings.log
