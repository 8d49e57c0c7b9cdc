# round 2, call M (1 GPU): split trace with a per-warp inline threshold (0 = always inline ... 33 = always defer)
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests -m gpu -x -q -k "variants or wavefront or exact_culling or golden" > gpurun_out/r02m/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02m/pytest.log
tail -3 gpurun_out/r02m/pytest.log
for t in 0 4 8 12 16 24 33; do
  echo "== inline_min $t" >> gpurun_out/r02m/probe.log
  for args in "c4 4 0" "c5 2 0" "c3 2 6"; do
    HAI719_MESH_INLINE_MIN=$t timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r02m/probe.log
  done
done
cat gpurun_out/r02m/probe.log
