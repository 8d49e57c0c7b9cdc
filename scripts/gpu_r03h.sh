# round 2, call 3h (1 GPU): trace kernel of scenes without meshes (k_wf_trace<.., 3>): parity, timings
mkdir -p gpurun_out/r03h
O=gpurun_out/r03h
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for args in "c2 16 0" "c2 64 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> $O/timings.log; done
cat $O/timings.log
