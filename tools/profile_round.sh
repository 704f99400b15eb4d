# Round profile pass: ncu launch list of the bench command + ncu --set full captures of the three DMMA kernels.
# Usage (GPU box): bash tools/profile_round.sh r01
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 300 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --structures 128 --no-cpu-baseline \
    > gpurun_out/b_ncu_$R.log 2>&1
for k in k_syrk_sk k_lrows_v3 k_xrows_v5; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$R \
      python tools/gpu_probe.py 96 --no-micro --no-simple > gpurun_out/ncu_${k}_$R.log 2>&1
done
ls -la gpurun_out/*_$R* | awk '{print $5, $9}'
