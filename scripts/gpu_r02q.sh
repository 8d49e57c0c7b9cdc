# round 2, call Q (1 GPU): normalized() shared (default) vs inlined everywhere vs more sharing
mkdir -p gpurun_out/r02q
bash scripts/gpu_ab.sh r02q/ab "default noinl0 noinl2" "c2 16 0" "c3 2 0" "c4 4 0" "c5 2 0"
cat gpurun_out/r02q/ab.log
