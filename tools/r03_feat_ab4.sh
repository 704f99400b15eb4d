run() { env "$@" python tools/cfg_probe.py $CFG $NS 2>&1 | grep "structures_per_s" | sed 's/.*"structures_per_s": \([0-9.]*\).*"features_G": \([0-9.]*\).*/structures_per_s \1 features_G \2/'; }
for CFG in 3 4; do
  if [ $CFG = 3 ]; then NS=24; else NS=16; fi
  for e in "PM_FEAT_NB=5 PM_FEAT_AT=1" "PM_FEAT_NB=2 PM_FEAT_AT=1" "PM_FEAT_NB=5 PM_FEAT_AT=2" "PM_FEAT_NB=2 PM_FEAT_AT=2"; do
    echo "== cfg $CFG $e"; run $e
  done
done
