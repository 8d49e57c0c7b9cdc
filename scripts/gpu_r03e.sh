# round 2, call 3e (1 GPU): lanes / issue per kernel of config 2 with and without direction-sorted windows
mkdir -p gpurun_out/r03e
HAI719_WF_SORT=1 bash scripts/gpu_issue.sh r03e/c2_sort1 c2 16
HAI719_WF_SORT=0 bash scripts/gpu_issue.sh r03e/c2_sort0 c2 16
ls -la gpurun_out/r03e
