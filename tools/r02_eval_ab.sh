# Eval A/B on the GPU box: default vs env switches.  Usage: bash tools/r02_eval_ab.sh <tag> [ENV=1 ...]
T=${1:-r02y}; shift
mkdir -p gpurun_out
python tools/eval_probe2.py 256 > gpurun_out/${T}_eval_default.log 2>&1; cat gpurun_out/${T}_eval_default.log
for e in "$@"; do
  env $e python tools/eval_probe2.py 256 > gpurun_out/${T}_eval_$e.log 2>&1; echo "== $e"; cat gpurun_out/${T}_eval_$e.log
done
