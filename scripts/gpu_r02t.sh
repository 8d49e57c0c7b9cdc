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
# config 1: kernel variants and CTAs per SM of k_render_paths, three decimals
for args in "c1 1 1 2 3 6"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r02t/c1.log; done
bash scripts/gpu_ab.sh r02t/ab_paths "default paths6 paths8" "c1 1 0"
cat gpurun_out/r02t/c1.log gpurun_out/r02t/ab_paths.log
