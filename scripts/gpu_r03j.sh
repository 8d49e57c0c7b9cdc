# round 2, call 3j (1 GPU): overflow queue packed into full batches: parity, on/off timings; source-level hot spots of bounce 1 (configs 4, 5, 2)
mkdir -p gpurun_out/r03j
O=gpurun_out/r03j
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for c in 1 0 2; do
  echo "== HAI719_WF_COMPACT=$c" >> $O/timings.log
  for args in "c5 2 0" "c3 2 6" "c4 4 0"; do
    HAI719_WF_COMPACT=$c timeout 600 python tools/variance_probe.py $args 2>&1 | grep "^upload 0" >> $O/timings.log
  done
done
cat $O/timings.log
LIB=hai719-raytracing_b200/lib/libhai719_rt.so
# config 4: kernels 4.. of the frame = bounce 1 (3 kernels per bounce + camera rays)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 3 --launch-count 2 -f -o /tmp/c4_l1 \
  python tools/profile_render.py --workload c4 --spp 1 --reps 1 --no-stats > $O/c4_l1_ncu.log 2>&1
python scripts/ncu_hotspots.py /tmp/c4_l1.ncu-rep $LIB --kernel "k_wf_trace" --index 0 --top 45 > $O/c4_l1_trace_a_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c4_l1.ncu-rep $LIB --kernel "k_wf_trace" --index 1 --top 45 > $O/c4_l1_trace_b_hotspots.txt 2>&1
# config 5: bounce 1 = kernels 4..7 (trace, classify, sample, sample-overflow)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 4 --launch-count 3 -f -o /tmp/c5_l1 \
  python tools/profile_render.py --workload c5 --spp 1 --reps 1 --no-stats > $O/c5_l1_ncu.log 2>&1
python scripts/ncu_hotspots.py /tmp/c5_l1.ncu-rep $LIB --kernel "k_wf_trace" --index 0 --top 45 > $O/c5_l1_trace_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c5_l1.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)1>" --index 0 --top 45 > $O/c5_l1_classify_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c5_l1.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)3>" --index 0 --top 45 > $O/c5_l1_sample_hotspots.txt 2>&1
# config 2: bounce 1 = kernels 3..5
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_.*ILb0 --launch-skip 3 --launch-count 3 -f -o /tmp/c2_l1 \
  python tools/profile_render.py --workload c2 --spp 16 --reps 1 --no-stats > $O/c2_l1_ncu.log 2>&1
python scripts/ncu_hotspots.py /tmp/c2_l1.ncu-rep $LIB --kernel "k_wf_trace" --index 0 --top 45 > $O/c2_l1_trace_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c2_l1.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)1>" --index 0 --top 45 > $O/c2_l1_classify_hotspots.txt 2>&1
python scripts/ncu_hotspots.py /tmp/c2_l1.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)3>" --index 0 --top 45 > $O/c2_l1_sample_hotspots.txt 2>&1
ls -la $O
