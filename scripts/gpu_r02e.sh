# round 2, call E (1 GPU): where the speculative shadow rounds spend their time (per-kernel durations, validation rates),
# larger fetch grabs, and the bulk-copy-staged analytic hierarchy A/B
mkdir -p gpurun_out/r02e
for args in "c5 2 1 2 3" "c3 2 1 2 3" "c2 4 1 2 3"; do timeout 300 python tools/spec_probe.py $args >> gpurun_out/r02e/spec_probe.log 2>&1; done
cat gpurun_out/r02e/spec_probe.log
for args in "c5 2 15728646 1048582 2097158" "c2 16 15728646 1048582 2097158"; do
  timeout 600 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/r02e/probe.log
done
cat gpurun_out/r02e/probe.log
bash scripts/gpu_times_wf.sh r02e/c5_spec1_times c5 2 1048582
bash scripts/gpu_times_wf.sh r02e/c5_spec0_times c5 2 15728646
bash scripts/gpu_ab.sh r02e/ab_stage "stage_abvh" "c2 16 0" "c4 4 0" "c5 2 0"
cat gpurun_out/r02e/ab_stage.log
