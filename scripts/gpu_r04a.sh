# round 2, call 4a (1 GPU): per-kernel figures of config 2 with and without the umbra short-cut in the classify kernel
mkdir -p gpurun_out/r04a
bash scripts/gpu_issue.sh r04a/c2_umbra1 c2 16
cp hai719-raytracing_b200/lib/libhai719_rt.so /tmp/keep.so; cp hai719-raytracing_b200/lib_alt/noumbra.so hai719-raytracing_b200/lib/libhai719_rt.so
bash scripts/gpu_issue.sh r04a/c2_umbra0 c2 16
cp /tmp/keep.so hai719-raytracing_b200/lib/libhai719_rt.so
