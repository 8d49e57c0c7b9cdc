set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest_gpu.log
for v in 0 4; do
 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r1_c2_v$v.json 2> gpurun_out/r1_c2_v$v.err
 timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r1_c3_v$v.json 2> gpurun_out/r1_c3_v$v.err
 timeout 300 python bench.py --workload c4 --spp 16 --steps 3 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r1_c4_v$v.json 2> gpurun_out/r1_c4_v$v.err
 timeout 300 python bench.py --workload c5 --spp 2 --steps 3 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r1_c5_v$v.json 2> gpurun_out/r1_c5_v$v.err
 timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/r1_c1_v$v.json 2> gpurun_out/r1_c1_v$v.err
done
timeout 300 python bench.py > gpurun_out/r1_default.json 2> gpurun_out/r1_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_reference.json 2> gpurun_out/r1_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_c2_ncu.log 2>&1
