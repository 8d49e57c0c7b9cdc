# round 2, call B (2 GPUs): the multi-GPU tests (rt_render_multi, IPC framebuffer), CLI --gpus 2, bench at N=2 with c4 and c5 sharded
mkdir -p gpurun_out/r02b
nvidia-smi -L > gpurun_out/r02b/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi or image_mode or ipc or shared or full_size" > gpurun_out/r02b/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/pytest_multi.log
for sc in 5 9; do
  timeout 120 hai719-raytracing_b200/bin/hai719_render --scene $sc --w 640 --h 360 --spp 8 --gpus 2 --out /tmp/multi_$sc.ppm --assets /root/repo/assets/_ref >> gpurun_out/r02b/cli_gpus2.log 2>&1; echo "scene $sc gpus 2 rc=$?" >> gpurun_out/r02b/cli_gpus2.log
  timeout 120 hai719-raytracing_b200/bin/hai719_render --scene $sc --w 640 --h 360 --spp 8 --out /tmp/single_$sc.ppm --assets /root/repo/assets/_ref >> gpurun_out/r02b/cli_gpus2.log 2>&1; echo "scene $sc gpus 1 rc=$?" >> gpurun_out/r02b/cli_gpus2.log
  cmp /tmp/multi_$sc.ppm /tmp/single_$sc.ppm >> gpurun_out/r02b/cli_gpus2.log 2>&1 && echo "scene $sc: 2-GPU file identical to 1-GPU file" >> gpurun_out/r02b/cli_gpus2.log
done
BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02b/bench_n2.json 2> gpurun_out/r02b/bench_n2.err; echo "rc=$?" >> gpurun_out/r02b/bench_n2.err
tail -3 gpurun_out/r02b/pytest_multi.log; tail -3 gpurun_out/r02b/cli_gpus2.log; grep -v "^\[W\|^W" gpurun_out/r02b/bench_n2.err | tail -8; head -c 400 gpurun_out/r02b/bench_n2.json
