# round 2, call 4b (1 GPU): approximate reciprocals / square roots inside the conservative candidate filters: parity, timings
mkdir -p gpurun_out/r04b
O=gpurun_out/r04b
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
cat $O/timings.log
