# round 2, call 4n (1 GPU): pond scene at 4 and 8 spp, state machine (3) against wavefront (6): where does the wavefront start to win?
mkdir -p gpurun_out/r04n
for spp in 4 8; do timeout 120 python tools/variance_probe.py c3 $spp 3 6 2>&1 | grep "^upload 0" | sed "s/^/spp $spp /" >> gpurun_out/r04n/c3_by_spp.log; done
cat gpurun_out/r04n/c3_by_spp.log
