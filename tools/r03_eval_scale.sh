for e in "A=1" "PM_EVAL_CHUNK_SCALE=0.7" "PM_EVAL_CHUNK_SCALE=0.5" "PM_EVAL_LANES=4 PM_EVAL_CHUNK_SCALE=0.65" "PM_EVAL_LANES=4"; do
  echo "== $e"; env $e python tools/eval_probe2.py 256 2>&1 | grep "atoms/s"; env $e python tools/eval_probe2.py 256 2>&1 | grep "atoms/s"
done
