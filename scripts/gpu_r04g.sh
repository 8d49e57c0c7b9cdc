# round 2, call 4g (1 GPU): flat box test compiled into the split trace kernel only: parity, timings of every config
mkdir -p gpurun_out/r04g
O=gpurun_out/r04g
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
HAI719_WF_FLAT=0 timeout 300 python tools/variance_probe.py c4 4 0 2>&1 | grep "^upload 0" >> $O/timings.log
cat $O/timings.log
