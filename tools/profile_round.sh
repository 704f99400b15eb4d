# Round profile pass (GPU box): ncu launch list of the bench command + ncu --set full captures of the hot kernels.
# Usage: bash tools/profile_round.sh r02
R=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 400 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --structures 128 --no-cpu-baseline --no-sub \
    > gpurun_out/b_ncu_$R.log 2>&1
echo "launch list rc=$?"
for k in k_syrk_sk2 k_lrows_v4n k_xrows_v6 k_pair_anlm k_features_v3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_$R \
      python tools/gpu_probe.py 80 --no-micro --no-simple > gpurun_out/ncu_${k}_$R.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page source --csv > gpurun_out/prof_${k}_$R.source.csv 2>/dev/null
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page raw --csv > gpurun_out/prof_${k}_$R.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${k}_$R.ncu-rep      # gpurun merges at most 64 MiB back: keep the two CSV pages only
done
for k in k_eval_features_lb k_eval_pairs_rc k_anlm_eval k_neighbor_cl_count; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$R \
      python tools/eval_probe2.py 64 > gpurun_out/ncu_${k}_$R.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page source --csv > gpurun_out/prof_${k}_$R.source.csv 2>/dev/null
  ncu -i gpurun_out/prof_${k}_$R.ncu-rep --page raw --csv > gpurun_out/prof_${k}_$R.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${k}_$R.ncu-rep
done
# eval: stage profile (one lane, host sync after every stage) and launch list
python tools/eval_probe2.py 256 > gpurun_out/eval_stages_$R.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 150 --csv \
   --log-file gpurun_out/eval_launches_$R.csv python tools/eval_probe2.py 128 > gpurun_out/eval_ncu_$R.log 2>&1
ls -la gpurun_out/*_$R* | awk '{print $5, $9}'
