# Round-2 GPU pass: new-kernel tests first (short timeouts: a hung mbarrier pipeline must not eat the box),
# then stage profiles under the A/B switches, then (optionally) the full GPU suite.
mkdir -p gpurun_out
T=${1:-r02a}
FULL=${2:-yes}
timeout 900 python -m pytest tests/test_gpu_round2.py -q > gpurun_out/${T}_pytest_round2.log 2>&1; echo "round2 tests rc=$?"
tail -15 gpurun_out/${T}_pytest_round2.log
while IFS= read -r cfg; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $( [ "$cfg" = default ] || echo $cfg ) timeout 300 python tools/gpu_probe.py 168 --no-micro --no-simple > gpurun_out/${T}_probe_${tag}.log 2>&1
  echo "== $cfg rc=$?"; grep -E "structures/s|lrows|xrows|syrk|anlm|Error" gpurun_out/${T}_probe_${tag}.log | tail -8
done < tools/r02_probe_cfgs.txt
if [ "$FULL" = yes ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
  tail -12 gpurun_out/${T}_pytest_gpu.log
fi
