# round 2, call 3v (1 GPU): large / sliver triangles split before the hierarchy build (HAI719_BVH_SPLIT = factor on the median box area): parity, sweep;
# and two build options on the specialised kernels (normalized() inlined; no streaming hints on the wavefront's records)
mkdir -p gpurun_out/r03v
O=gpurun_out/r03v
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for f in 0 2 4 8 16 32; do
  echo "== HAI719_BVH_SPLIT=$f" >> $O/split.log
  for args in "c3 2 0" "c4 4 0" "c5 2 0"; do
    HAI719_BVH_SPLIT=$f timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/split.log
  done
done
cat $O/split.log
bash scripts/gpu_ab.sh r03v/ab_build "noinl0 stream0" "c2 16 0" "c2 64 0" "c5 2 0"
cat $O/ab_build.log
