# round 2, call C (1 GPU): parity of the constant-bank scene + shared-memory traversal stack build, then A/B of the variants
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02c/pytest.log
tail -3 gpurun_out/r02c/pytest.log
bash scripts/gpu_ab.sh r02c/ab "base const const_sm12 const_sm16 const_sm24 const_sm16_6cta" "c2 16 0" "c3 2 0" "c4 4 0" "c5 2 0" "c1 1 0"
cat gpurun_out/r02c/ab.log | grep -v "^upload 0"
