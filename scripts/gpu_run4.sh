mkdir -p gpurun_out
bash scripts/gpu_ab.sh r6_ab "n_cur o_noreach p_noww q_neither" "c3 2 0" "c4 2 0" "c5 2 0" "c2 8 0"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r6_pytest.log
bash scripts/gpu_prof.sh r6_prof_c2 c2 8
