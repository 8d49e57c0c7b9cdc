mkdir -p gpurun_out
rm -f gpurun_out/i_ab.log gpurun_out/i_chunk.log
bash scripts/gpu_ab.sh i_ab "nopf pf" "c2 32 6" "c5 2 6" "c4 4 6"
for l in 25 26; do echo "chunk log2 $l" >> gpurun_out/i_chunk.log; HAI719_CHUNK_LOG2=$l timeout 300 python tools/variance_probe.py c2 64 6 2>&1 | grep -v "^upload [12]" >> gpurun_out/i_chunk.log; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants_agree or culling_at_scale or output_stage or chunked" > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log
