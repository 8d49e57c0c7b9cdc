# round 2, call G (1 GPU): queue-position records + block reservation + no-light trace: memcheck on a small render, parity, timings
mkdir -p gpurun_out/r02g
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_render.py > gpurun_out/r02g/sanitize.log 2>&1; echo "rc=$?" >> gpurun_out/r02g/sanitize.log
tail -4 gpurun_out/r02g/sanitize.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02g/pytest.log
tail -3 gpurun_out/r02g/pytest.log
for args in "c2 16 0" "c3 2 0 6" "c4 4 0" "c5 2 0" "c1 1 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02g/probe.log
done
cat gpurun_out/r02g/probe.log
bash scripts/gpu_ab.sh r02g/ab_paths "paths6 paths8" "c1 1 0"
cat gpurun_out/r02g/ab_paths.log
