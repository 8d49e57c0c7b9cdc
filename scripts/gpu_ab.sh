# A/B of alternate builds of the CUDA library: bash scripts/gpu_ab.sh TAG "name1 name2 ..." "probe args"
TAG=$1; NAMES=$2; shift 2
mkdir -p gpurun_out
cp hai719-raytracing_b200/lib/libhai719_rt.so /tmp/rt_keep.so
for n in $NAMES; do
  cp hai719-raytracing_b200/lib_alt/$n.so hai719-raytracing_b200/lib/libhai719_rt.so
  echo "=== $n" >> gpurun_out/${TAG}.log
  for args in "$@"; do
    timeout 300 python tools/variance_probe.py $args 2>&1 | grep -v "^upload [12]" >> gpurun_out/${TAG}.log
  done
done
cp /tmp/rt_keep.so hai719-raytracing_b200/lib/libhai719_rt.so
