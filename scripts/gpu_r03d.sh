# round 2, call 3d (1 GPU): live-queue windows sorted by ray direction (wf_next_batch_sorted): parity, on/off timings, lanes per kernel
mkdir -p gpurun_out/r03d
O=gpurun_out/r03d
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for sort in 1 0; do
  echo "== HAI719_WF_SORT=$sort" >> $O/timings.log
  for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0"; do
    HAI719_WF_SORT=$sort timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log
  done
done
for grab in 2 4 8; do
  echo "== HAI719_WF_SORT=1 HAI719_WF_GRAB=$grab (mesh scenes)" >> $O/timings.log
  for args in "c4 4 0" "c5 2 0"; do
    HAI719_WF_SORT=1 HAI719_WF_GRAB=$grab timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log
  done
done
cat $O/timings.log
