# round 2, call A (1 GPU): all GPU tests, smoke, the new bench (every config at full size), reference arm
mkdir -p gpurun_out/r02a
nvidia-smi -L > gpurun_out/r02a/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02a/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a/smoke.log 2>&1
BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/r02a/bench_default.json 2> gpurun_out/r02a/bench_default.err; echo "rc=$?" >> gpurun_out/r02a/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a/bench_reference.json 2> gpurun_out/r02a/bench_reference.err
tail -3 gpurun_out/r02a/pytest.log; cat gpurun_out/r02a/smoke.log | tail -2; tail -5 gpurun_out/r02a/bench_default.err; head -c 600 gpurun_out/r02a/bench_default.json
