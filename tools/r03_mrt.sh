for CFG in 3 4; do
  if [ $CFG = 3 ]; then NS=24; else NS=16; fi
  for mrt in 0 8 9 12 16; do
    echo "== cfg $CFG mrt $mrt"; PM_LROWS_MRT=$mrt python tools/cfg_probe.py $CFG $NS 2>&1 | grep structures_per_s | sed "s/.*\"structures_per_s\": \([0-9.]*\).*\"lrows\": \([0-9.]*\).*/structures_per_s \1 lrows \2/"
  done
done
