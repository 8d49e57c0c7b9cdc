mkdir -p gpurun_out
rm -f gpurun_out/h_ab.log gpurun_out/h_chunk.log
bash scripts/gpu_ab.sh h_ab "cur lc4 pf" "c3 8 3" "c5 2 6" "c2 8 6" "c4 2 6"
for l in 22 23 24 25; do echo "chunk log2 $l" >> gpurun_out/h_chunk.log; HAI719_CHUNK_LOG2=$l timeout 300 python tools/variance_probe.py c2 32 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/h_chunk.log; done
