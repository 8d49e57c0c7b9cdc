# round 2, call D (1 GPU): parity with speculative shadow samples, then timings with 0 / 1 / 2 / 3 / 6 rounds
mkdir -p gpurun_out/r02d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02d/pytest.log
tail -3 gpurun_out/r02d/pytest.log
for args in "c3 2 3 15728646 1048582 2097158 3145734 6291462" "c5 2 15728646 1048582 2097158 3145734 6291462" "c2 16 15728646 3145734" "c4 4 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02d/probe.log
done
cat gpurun_out/r02d/probe.log
