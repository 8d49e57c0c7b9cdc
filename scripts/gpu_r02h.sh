# round 2, call H (1 GPU): parity of the slimmer queue code; no-light trace on/off on config 4; traffic + issue per kernel
mkdir -p gpurun_out/r02h
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02h/pytest.log
tail -3 gpurun_out/r02h/pytest.log
for args in "c2 16 0" "c4 4 0 536870918" "c5 2 0" "c3 2 0 6"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02h/probe.log
done
cat gpurun_out/r02h/probe.log
bash scripts/gpu_issue.sh r02h/c2_issue c2 16
bash scripts/gpu_issue.sh r02h/c4_issue c4 4
bash scripts/gpu_issue.sh r02h/c4_issue_nolight c4 4 536870918
bash scripts/gpu_issue.sh r02h/c5_issue c5 2
for c in c2_issue c4_issue c4_issue_nolight c5_issue; do python scripts/ncu_issue_summary.py gpurun_out/r02h/$c.csv; done
