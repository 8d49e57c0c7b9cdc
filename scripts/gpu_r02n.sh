# round 2, call N (1 GPU): candidate list out of the path state; light-stage helpers inlined or not, 8x64 or 6x80 registers
mkdir -p gpurun_out/r02n
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02n/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02n/pytest.log
tail -3 gpurun_out/r02n/pytest.log
bash scripts/gpu_ab.sh r02n/ab "default inl inl6 inls inlb minb6" "c2 16 0" "c3 2 0 6" "c4 4 0" "c5 2 0"
cat gpurun_out/r02n/ab.log
