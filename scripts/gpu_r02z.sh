# round 2, call Z (1 GPU): where the fixed cost of a small chunk goes: per-kernel times of config 2 at 4 / 8 / 16 spp
mkdir -p gpurun_out/r02z
for spp in 4 8 16; do bash scripts/gpu_issue.sh r02z/c2_${spp}spp c2 $spp; done
for spp in 4 8 16; do timeout 300 python tools/variance_probe.py c2 $spp 0 2>&1 | grep "^upload 1" >> gpurun_out/r02z/probe.log; done
cat gpurun_out/r02z/probe.log
