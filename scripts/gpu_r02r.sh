# round 2, call R (1 GPU): 4-wide mesh hierarchy (RT_OPT_BVH4) against the binary one
mkdir -p gpurun_out/r02r
cp hai719-raytracing_b200/lib/libhai719_rt.so /tmp/keep.so
cp hai719-raytracing_b200/lib_alt/bvh4.so hai719-raytracing_b200/lib/libhai719_rt.so
timeout 900 python -m pytest tests -m gpu -x -q -k "variants or exact_culling or wavefront or golden or degenerate" > gpurun_out/r02r/pytest_bvh4.log 2>&1; echo "rc=$?" >> gpurun_out/r02r/pytest_bvh4.log
tail -3 gpurun_out/r02r/pytest_bvh4.log
cp /tmp/keep.so hai719-raytracing_b200/lib/libhai719_rt.so
bash scripts/gpu_ab.sh r02r/ab "default bvh4" "c3 2 0 6" "c4 4 0" "c5 2 0"
cat gpurun_out/r02r/ab.log
