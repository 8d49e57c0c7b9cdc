# round 2, call 3r (1 GPU): prepadded analytic hierarchy (intersect_lc<.., PREPAD>): parity, A/B against the per-ray padding
mkdir -p gpurun_out/r03r
O=gpurun_out/r03r
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
echo "=== default (prepadded)" >> $O/timings.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 6"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
bash scripts/gpu_ab.sh r03r/timings "noprepad" "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 6"
cat $O/timings.log
