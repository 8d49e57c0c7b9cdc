# round 2, call 4e (1 GPU): flat box test over <= 32 analytic primitives (pool scene: 30) instead of the hierarchy walk: parity, A/B
mkdir -p gpurun_out/r04e
O=gpurun_out/r04e
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for f in 2 0 1; do
  echo "== HAI719_WF_FLAT=$f" >> $O/timings.log
  HAI719_WF_FLAT=$f timeout 600 python tools/variance_probe.py c4 4 0 2>&1 | grep "^upload 0" >> $O/timings.log
done
cat $O/timings.log
