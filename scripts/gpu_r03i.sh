# round 2, call 3i (1 GPU): lanes / issue per kernel, current build
mkdir -p gpurun_out/r03i
bash scripts/gpu_issue.sh r03i/c2 c2 16
bash scripts/gpu_issue.sh r03i/c5 c5 2
bash scripts/gpu_issue.sh r03i/c4 c4 4
