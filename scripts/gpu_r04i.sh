# round 2, call 4i (1 GPU): any-hit mesh walk for the shadow samples of the overflow queue: parity, A/B
mkdir -p gpurun_out/r04i
O=gpurun_out/r04i
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
echo "=== default (any-hit)" >> $O/timings.log
for args in "c5 2 0" "c3 2 6" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log; done
bash scripts/gpu_ab.sh r04i/timings "noanyhit" "c5 2 0" "c3 2 6"
cat $O/timings.log
