# Eval part of the round profile (GPU box).  Usage: bash tools/profile_eval.sh r02
R=${1:-r02}
mkdir -p gpurun_out
for k in k_eval_features_lb k_eval_pairs_rc k_anlm_eval k_neighbor_cl_count k_neighbor_cl_fill; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$R \
      python tools/eval_probe2.py 64 > gpurun_out/ncu_${k}_$R.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page source --csv > gpurun_out/prof_${k}_$R.source.csv 2>/dev/null
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page raw --csv > gpurun_out/prof_${k}_$R.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${k}_$R.ncu-rep
done
python tools/eval_probe2.py 256 > gpurun_out/eval_stages_$R.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 150 --csv \
   --log-file gpurun_out/eval_launches_$R.csv python tools/eval_probe2.py 128 > gpurun_out/eval_ncu_$R.log 2>&1
cat gpurun_out/eval_stages_$R.log
