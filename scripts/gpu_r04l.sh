# round 2, call 4l (1 GPU): the bench line of the final tree + the issue captures that feed profiles/latest.json
mkdir -p gpurun_out/r04l
O=gpurun_out/r04l
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
BENCH_DEBUG=1 timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
bash scripts/gpu_issue.sh r04l/c2_issue c2 16
bash scripts/gpu_issue.sh r04l/c3_issue c3 2
bash scripts/gpu_issue.sh r04l/c4_issue c4 4
bash scripts/gpu_issue.sh r04l/c5_issue c5 2
bash scripts/gpu_issue.sh r04l/c1_issue c1 1
tail -2 $O/smoke.log; grep -v "^\[W" $O/bench_default.err | tail -6; head -c 200 $O/bench_default.json
