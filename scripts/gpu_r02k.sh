# round 2, call K (1 GPU): split trace (analytic phase + mesh-walk kernel) and the lean scatter kernel for scenes without lights
mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02k/pytest.log
tail -3 gpurun_out/r02k/pytest.log
# variants: 0 auto; 268435462 = 6|bit28 (no split); 134217734 = 6|bit27 (general light kernel); 402653190 = both off
for args in "c4 4 0 268435462 134217734 402653190" "c5 2 0 268435462" "c3 2 0 6 268435462" "c2 16 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02k/probe.log
done
cat gpurun_out/r02k/probe.log
bash scripts/gpu_issue.sh r02k/c4_issue c4 4
bash scripts/gpu_issue.sh r02k/c5_issue c5 2
for c in c4_issue c5_issue; do python scripts/ncu_issue_summary.py gpurun_out/r02k/$c.csv; done
