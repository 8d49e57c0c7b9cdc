# round-end validation on one GPU (everything the driver will run, plus the ncu evidence bench.py cites):
#   bash scripts/gpu_final.sh TAG
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
BENCH_DEBUG=1 timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
# issue-side utilisation and DRAM traffic of every kernel of one frame (feeds profiles/latest.json)
bash scripts/gpu_issue.sh $TAG/c2_issue c2 16
bash scripts/gpu_issue.sh $TAG/c3_issue c3 2
bash scripts/gpu_issue.sh $TAG/c4_issue c4 4
bash scripts/gpu_issue.sh $TAG/c5_issue c5 2
bash scripts/gpu_issue.sh $TAG/c1_issue c1 1
# launch list of the bench command itself (headline only: per-launch times under ncu are cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scenes none > $O/bench_under_ncu.log 2>&1
# ncu --set full of the first wavefront kernels of config 2 (one chunk) and of config 4
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_wf_ --launch-count 6 -f -o $O/c2_wf python tools/profile_render.py --workload c2 --spp 16 --reps 1 --no-stats > $O/c2_wf_ncu.log 2>&1
# hot spots by source line of the three kernels of bounce 0 (scripts/ncu_hotspots.py joins the report with the library's line table);
# gpurun_out/ may not exceed 64 MiB: the report itself travels xz-compressed
LIB=hai719-raytracing_b200/lib/libhai719_rt.so
python scripts/ncu_hotspots.py $O/c2_wf.ncu-rep $LIB --kernel k_wf_trace --index 0 --top 40 > $O/c2_trace_b0_hotspots.txt 2>&1
python scripts/ncu_hotspots.py $O/c2_wf.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)1" --index 0 --top 40 > $O/c2_classify_b0_hotspots.txt 2>&1
python scripts/ncu_hotspots.py $O/c2_wf.ncu-rep $LIB --kernel "k_wf_light<(bool)0, (bool)1, (int)3" --index 0 --top 40 > $O/c2_sample_b0_hotspots.txt 2>&1
python scripts/ncu_summary.py $O/c2_wf.ncu-rep > $O/c2_wf_ncu_summary.txt 2>&1
ncu -i $O/c2_wf.ncu-rep --page raw --csv > $O/c2_wf_ncu_raw.csv 2>/dev/null
(cd $O && for f in *.ncu-rep; do xz -T8 -6 $f; done)
tail -3 $O/pytest.log; tail -2 $O/smoke.log; grep -v "^\[W" $O/bench_default.err | tail -6; head -c 300 $O/bench_default.json; echo; du -sh gpurun_out
