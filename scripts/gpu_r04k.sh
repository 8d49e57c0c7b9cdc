# round 2, call 4k (1 GPU): the wavefront as the automatic choice for every scene with the analytic hierarchy (pond scene included): parity, timings
mkdir -p gpurun_out/r04k
O=gpurun_out/r04k
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c3 16 0" "c3 2 0" "c5 2 0" "c2 64 0" "c4 4 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
cat $O/timings.log
