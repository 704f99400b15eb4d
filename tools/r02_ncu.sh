# ncu --set full captures of the named kernels (one launch each) on a config-2 chunk of N structures.
# Usage (GPU box): bash tools/r02_ncu.sh <tag> <n structures> <skip launches> <kernel regex> [<kernel regex> ...]
mkdir -p gpurun_out
T=$1; N=$2; S=$3; shift; shift; shift
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $S -c 1 -f -o gpurun_out/${T}_$k \
      python tools/gpu_probe.py $N --no-micro --no-simple > gpurun_out/${T}_ncu_$k.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/${T}_$k.ncu-rep --page source --csv > gpurun_out/${T}_$k.source.csv 2>/dev/null
done
ls -la gpurun_out/ | awk '{print $5, $9}' | grep $T
