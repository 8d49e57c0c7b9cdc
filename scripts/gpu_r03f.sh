# round 2, call 3f (1 GPU): classify kernel specialised to the cone walk (intersect_lc<.., COLLECT>): parity, A/B against the general build
mkdir -p gpurun_out/r03f
O=gpurun_out/r03f
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
bash scripts/gpu_ab.sh r03f/ab_lc_collect "lc_nocollect" "c2 16 0" "c2 64 0" "c5 2 0" "c3 2 6"
echo "=== default (collect-only classify)" >> $O/ab_lc_collect.log
for args in "c2 16 0" "c2 64 0" "c5 2 0" "c3 2 6"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> $O/ab_lc_collect.log; done
cat $O/ab_lc_collect.log
