# Eval (config-5 shape) stage profile + launch list on the GPU box.  Usage: bash tools/r02_eval_stages.sh <tag>
T=${1:-r02x}
mkdir -p gpurun_out
python tools/eval_probe2.py 256 > gpurun_out/${T}_eval_stages.log 2>&1
PM_DEBUG_TIMING=1 python tools/eval_probe2.py 256 > gpurun_out/${T}_eval_debug.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv \
   --log-file gpurun_out/${T}_eval_launches.csv python tools/eval_probe2.py 128 > gpurun_out/${T}_eval_ncu.log 2>&1
cat gpurun_out/${T}_eval_stages.log
tail -30 gpurun_out/${T}_eval_debug.log
