# round 2, call F (1 GPU): ncu --set full of config 3's state-machine kernel and of the first wavefront kernels of configs 4 and 2;
# issue-side utilisation of every kernel of one frame of configs 2, 3, 4; then parity + e2e probes of the image-cache build
mkdir -p gpurun_out/r02f
bash scripts/gpu_prof.sh r02f/c3_regen c3 1 0
bash scripts/gpu_prof_wf.sh r02f/c4_wf c4 2 0 4
bash scripts/gpu_prof_wf.sh r02f/c2_wf c2 16 0 6
bash scripts/gpu_issue.sh r02f/c2_issue c2 16
bash scripts/gpu_issue.sh r02f/c3_issue c3 2
bash scripts/gpu_issue.sh r02f/c4_issue c4 4
(cd gpurun_out/r02f && for f in *.ncu-rep; do xz -T8 -3 $f; done; ls -la)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f/pytest.log
tail -3 gpurun_out/r02f/pytest.log
for w in c1 c2 c4; do timeout 300 python tools/e2e_probe.py $w 20 >> gpurun_out/r02f/e2e_probe.log 2>&1; done
cat gpurun_out/r02f/e2e_probe.log
du -sh gpurun_out
