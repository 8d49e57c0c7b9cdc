# round 2, call W (1 GPU): equal chunks, 64 Mi-path chunks for scenes without meshes; parity; C2 at 8 spp (the size of an 8-GPU shard)
mkdir -p gpurun_out/r02w
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02w/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02w/pytest.log
tail -3 gpurun_out/r02w/pytest.log
for args in "c2 64 0" "c2 8 0" "c2 16 0" "c5 4 0" "c4 8 0" "c3 4 0"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 1" >> gpurun_out/r02w/probe.log
done
for l in 25 26; do echo "== chunk log2 $l" >> gpurun_out/r02w/probe.log; HAI719_CHUNK_LOG2=$l timeout 600 python tools/variance_probe.py c2 64 0 2>&1 | grep "^upload 1" >> gpurun_out/r02w/probe.log; done
cat gpurun_out/r02w/probe.log
