# round 2, call 3q (1 GPU): sample kernel compiled per queue (list / overflow): parity, timings
mkdir -p gpurun_out/r03q
O=gpurun_out/r03q
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c5 2 0" "c3 2 6" "c3 2 0" "c2 16 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
cat $O/timings.log
