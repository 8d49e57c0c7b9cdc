# round 2, call 3x (1 GPU): CTAs per SM of the no-mesh kernel instantiations (sample 7 / 6, trace 7, classify 7 against 8 everywhere): A/B on config 2
mkdir -p gpurun_out/r03x
bash scripts/gpu_ab.sh r03x/ab_minb "s7 s6 t7 c7" "c2 16 0" "c2 64 0"
echo "=== default" >> gpurun_out/r03x/ab_minb.log
for args in "c2 16 0" "c2 64 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r03x/ab_minb.log; done
cat gpurun_out/r03x/ab_minb.log
