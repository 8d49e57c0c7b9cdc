# progressive accumulation / preview: the GPU suite, then the default bench (kernels gained a sample_base argument) and the CLI preview
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err
( cd gpurun_out && timeout 300 ../hai719-raytracing_b200/bin/hai719_render --scene 5 --assets ../assets/_ref --w 1920 --h 1080 --spp 4 --preview 6 --orbit 40 --out g_preview > g_preview.log 2>&1; echo "rc=$?" >> g_preview.log; ls -la g_preview.*.ppm >> g_preview.log; rm -f g_preview.*.ppm )
