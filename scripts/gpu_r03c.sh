# round 2, call 3c (1 GPU): kernel variant 7 (state machine with the resumable mesh walk): parity, then threshold sweep
mkdir -p gpurun_out/r03c
O=gpurun_out/r03c
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or culling_at_scale" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
v() { echo $(( 7 | ($1 << 20) | ($2 << 8) )); }
VARS="3 7"
for T in 4 8 12 16 20 24 28; do for R in 1 4 16; do VARS="$VARS $(v $T $R)"; done; done
timeout 900 python tools/variance_probe.py c3 2 $VARS 2>&1 | grep "^upload 0" > $O/c3_sweep.log
cat $O/c3_sweep.log
timeout 600 python tools/variance_probe.py c5 2 0 3 7 $(v 8 4) $(v 16 4) $(v 24 4) 2>&1 | grep "^upload 0" > $O/c5_sweep.log
cat $O/c5_sweep.log
timeout 600 python tools/variance_probe.py c4 4 0 3 7 $(v 8 4) $(v 16 4) $(v 24 4) 2>&1 | grep "^upload 0" > $O/c4_sweep.log
cat $O/c4_sweep.log
