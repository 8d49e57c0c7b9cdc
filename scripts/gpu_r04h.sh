# round 2, call 4h (1 GPU): capacity of the candidate-triangle list (8 / 16 / 24 / 32) on the final kernels: A/B on the lit mesh scenes
mkdir -p gpurun_out/r04h
bash scripts/gpu_ab.sh r04h/ab_maxc "lc8 lc24 lc32" "c5 2 0" "c3 2 6"
echo "=== default (16)" >> gpurun_out/r04h/ab_maxc.log
for args in "c5 2 0" "c3 2 6"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r04h/ab_maxc.log; done
cat gpurun_out/r04h/ab_maxc.log
