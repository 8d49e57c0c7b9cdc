# round 2, call 3a (1 GPU): two lanes (streams) per frame in the wavefront: parity, then on/off timings
mkdir -p gpurun_out/r03a
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03a/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r03a/pytest.log
tail -3 gpurun_out/r03a/pytest.log
for lanes in 2 1; do
  echo "== lanes $lanes" >> gpurun_out/r03a/probe.log
  for args in "c2 64 0" "c2 8 0" "c2 2 0" "c4 8 0" "c5 4 0" "c3 4 6"; do
    HAI719_LANES=$lanes timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r03a/probe.log
  done
done
cat gpurun_out/r03a/probe.log
