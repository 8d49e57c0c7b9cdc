# strong-scaling check on one box: bash scripts/gpu_scale.sh "1 2 4 8"
mkdir -p gpurun_out
for n in $1; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
done
