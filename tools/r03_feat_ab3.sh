for e in "A=1" "PM_FEAT_NB=2" "PM_FEAT_NB=1" "PM_FEAT_V3=1"; do
  echo "== $e"; env $e PM_DEBUG_TABLES=1 python tools/cfg_probe.py 3 24 2>&1 | grep "structures_per_s\|radial-0" | sed 's/.*"structures_per_s": \([0-9.]*\).*"features_G": \([0-9.]*\).*/structures_per_s \1 features_G \2/' | head -4
done
echo "== config 4"; python tools/cfg_probe.py 4 16 2>&1 | grep "structures_per_s" | sed 's/.*"structures_per_s": \([0-9.]*\).*"features_G": \([0-9.]*\).*/structures_per_s \1 features_G \2/'
PM_FEAT_V3=1 python tools/cfg_probe.py 4 16 2>&1 | grep "structures_per_s" | sed 's/.*"structures_per_s": \([0-9.]*\).*"features_G": \([0-9.]*\).*/structures_per_s \1 features_G \2/'
