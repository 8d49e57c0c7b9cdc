mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r10_pytest.log
bash scripts/gpu_prof_wf.sh r10_prof_wf_c2 c2 8
