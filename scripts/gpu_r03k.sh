# round 2, call 3k (1 GPU): linear analytic loops from bounce 1 on for scenes with few analytic primitives: parity, on/off timings
mkdir -p gpurun_out/r03k
O=gpurun_out/r03k
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for c in 1 0; do
  echo "== HAI719_WF_LINEAR=$c" >> $O/timings.log
  for args in "c4 4 0" "c3 2 6" "c2 16 0" "c5 2 0"; do
    HAI719_WF_LINEAR=$c timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log
  done
done
cat $O/timings.log
