# round 2, call 3g (1 GPU): sample kernel of one-light scenes compiled for the sample tests only (k_wf_light<.., 3>): parity, timings
mkdir -p gpurun_out/r03g
O=gpurun_out/r03g
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c2 16 0" "c2 64 0" "c5 2 0" "c3 2 6" "c4 4 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> $O/timings.log; done
cat $O/timings.log
