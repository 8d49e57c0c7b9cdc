# round 2, call V (1 GPU): speculative samples for the overflow queue (parity, rounds 0..3 on c3 wavefront / c5), then call T's A/Bs
mkdir -p gpurun_out/r02v
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02v/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v/pytest.log
tail -3 gpurun_out/r02v/pytest.log
# 15728646 = 6 | 15<<20 (off), 1048582 / 2097158 / 3145734 / 4194310 = 1 / 2 / 3 / 4 rounds
for args in "c3 2 3 15728646 1048582 2097158 3145734 4194310" "c5 2 15728646 1048582 2097158 3145734"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r02v/spec.log
done
cat gpurun_out/r02v/spec.log
bash scripts/gpu_r02t.sh
