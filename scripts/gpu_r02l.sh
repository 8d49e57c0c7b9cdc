# round 2, call L (1 GPU): ray by value into the out-of-line reachability functions (path state of the trace kernels back in registers)
mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02l/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02l/pytest.log
tail -3 gpurun_out/r02l/pytest.log
for args in "c2 16 0" "c4 4 0 268435462" "c5 2 0 268435462" "c3 2 0 6 268435462" "c1 1 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02l/probe.log
done
cat gpurun_out/r02l/probe.log
bash scripts/gpu_issue.sh r02l/c2_issue c2 16
bash scripts/gpu_issue.sh r02l/c3_issue c3 2
for c in c2_issue c3_issue; do python scripts/ncu_issue_summary.py gpurun_out/r02l/$c.csv; done
