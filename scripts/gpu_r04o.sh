# round 2, call 4o (1 GPU): last check of the final tree: the GPU suite (size rule for the pond scene's kernel in place)
mkdir -p gpurun_out/r04o
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r04o/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r04o/pytest.log
tail -3 gpurun_out/r04o/pytest.log
