# round 2, call O (1 GPU): light-stage helpers inlined in the wavefront only; new binding / image-cache tests
mkdir -p gpurun_out/r02o
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02o/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02o/pytest.log
tail -3 gpurun_out/r02o/pytest.log
for args in "c2 16 0" "c3 2 0 6" "c4 4 0" "c5 2 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02o/probe.log
done
cat gpurun_out/r02o/probe.log
