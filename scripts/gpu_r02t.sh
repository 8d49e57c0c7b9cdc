# round 2, call T (1 GPU): candidate-list capacity 8 / 16 / 32; chunk size 32 Mi vs 64 Mi paths
mkdir -p gpurun_out/r02t
bash scripts/gpu_ab.sh r02t/ab_lc "default lc8 lc32" "c3 2 6" "c5 2 0" "c4 4 0"
cat gpurun_out/r02t/ab_lc.log
for l in 25 26; do
  echo "== chunk log2 $l" >> gpurun_out/r02t/chunk.log
  for args in "c2 64 0" "c5 4 0" "c4 8 0"; do
    HAI719_CHUNK_LOG2=$l timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r02t/chunk.log
  done
done
cat gpurun_out/r02t/chunk.log
