# strong scaling of every config on N GPUs of one box: bash scripts/gpu_scale8.sh N TAG
N=$1; TAG=$2
mkdir -p gpurun_out/$TAG
nvidia-smi -L > gpurun_out/$TAG/gpus.txt 2>&1
BENCH_DEBUG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err; echo "rc=$?" >> gpurun_out/$TAG/bench_n$N.err
grep -v "^\[W\|^W" gpurun_out/$TAG/bench_n$N.err | tail -8; head -c 300 gpurun_out/$TAG/bench_n$N.json
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi" > gpurun_out/$TAG/pytest_multi.log 2>&1; tail -2 gpurun_out/$TAG/pytest_multi.log
