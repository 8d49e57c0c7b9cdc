# round 2, call 3l (1 GPU): classify in two kernels (mesh part of the candidate collection in k_wf_light<.., 4>): parity, on/off timings
mkdir -p gpurun_out/r03l
O=gpurun_out/r03l
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for c in 1 0; do
  echo "== HAI719_WF_CSPLIT=$c" >> $O/timings.log
  for args in "c5 2 0" "c3 2 6" "c2 16 0"; do
    HAI719_WF_CSPLIT=$c timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log
  done
done
cat $O/timings.log
