# ncu --set full captures of the eval kernels on a 64-structure config-5 batch.  Usage: bash tools/r02_ncu_eval.sh <tag> <kernel regex>...
mkdir -p gpurun_out
T=$1; shift
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${T}_$k \
      python tools/eval_probe2.py 64 > gpurun_out/${T}_ncu_$k.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/${T}_$k.ncu-rep --page source --csv > gpurun_out/${T}_$k.source.csv 2>/dev/null
done
