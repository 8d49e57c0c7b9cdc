# round 2, call 3n (1 GPU): box / cone tests as one FMA per plane (RT_OPT_BOXFMA, RT_OPT_CONEFMA) on the specialised kernels: A/B
mkdir -p gpurun_out/r03n
bash scripts/gpu_ab.sh r03n/ab_fma "fma boxfma" "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"
echo "=== default" >> gpurun_out/r03n/ab_fma.log
for args in "c2 16 0" "c2 64 0" "c4 4 0" "c5 2 0" "c3 2 0"; do timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r03n/ab_fma.log; done
cat gpurun_out/r03n/ab_fma.log
