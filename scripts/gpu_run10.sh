mkdir -p gpurun_out
rm -f gpurun_out/r15_sweep.log
for wl in "c2 8 6" "c5 2 6" "c4 2 6" "c3 2 6 3"; do
timeout 600 python tools/variance_probe.py $wl 2>&1 | grep -v "^upload [12]" >> gpurun_out/r15_sweep.log
done
