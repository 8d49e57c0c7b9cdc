# round 2, call 4j (1 GPU): pond scene at its full 16 spp, state machine (auto) against the wavefront
mkdir -p gpurun_out/r04j
timeout 600 python tools/variance_probe.py c3 16 0 6 2>&1 | grep "^upload [01]" > gpurun_out/r04j/c3_full.log
cat gpurun_out/r04j/c3_full.log
