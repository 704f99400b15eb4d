# Round-2 GPU pass: new-kernel tests first (short timeouts: a hung mbarrier pipeline must not eat the box),
# then stage profiles of both front-end generations / both SYRK kernels, then the full GPU suite.
mkdir -p gpurun_out
T=${1:-r02a}
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q > gpurun_out/${T}_pytest_round2.log 2>&1; echo "round2 tests rc=$?"
tail -5 gpurun_out/${T}_pytest_round2.log
for cfg in "default" "PM_FRONT_V1=1" "PM_SYRK_V1=1" "PM_FRONT_V1=1 PM_SYRK_V1=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $( [ "$cfg" = default ] || echo $cfg ) timeout 300 python tools/gpu_probe.py 168 --no-micro --no-simple > gpurun_out/${T}_probe_${tag}.log 2>&1
  echo "== $cfg rc=$?"; cat gpurun_out/${T}_probe_${tag}.log | tail -12
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -8 gpurun_out/${T}_pytest_gpu.log
