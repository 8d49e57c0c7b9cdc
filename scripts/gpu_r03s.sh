# round 2, call 3s (1 GPU): quadratic term of the per-ray box padding from the ball of the sphere centres: parity, timings
mkdir -p gpurun_out/r03s
O=gpurun_out/r03s
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
cat $O/timings.log
