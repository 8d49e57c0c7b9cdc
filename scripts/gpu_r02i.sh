# round 2, call I (1 GPU): parity + timings of the queue-position build after the register-pressure fixes; issue/traffic of c2 and c4
mkdir -p gpurun_out/r02i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02i/pytest.log
tail -3 gpurun_out/r02i/pytest.log
for args in "c2 16 0" "c4 4 0 536870918" "c5 2 0" "c3 2 0 6"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02i/probe.log
done
cat gpurun_out/r02i/probe.log
bash scripts/gpu_issue.sh r02i/c2_issue c2 16
bash scripts/gpu_issue.sh r02i/c4_issue c4 4
for c in c2_issue c4_issue; do python scripts/ncu_issue_summary.py gpurun_out/r02i/$c.csv; done
