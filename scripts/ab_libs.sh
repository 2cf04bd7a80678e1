# usage: ab_libs.sh libA libB ... ; alternates quick_rate.py over the given libraries (development aid)
cd /root/repo
for rep in 1 2 3; do
for lib in "$@"; do
  for G in ${GROUPS_:-2 4}; do
    DPHLM_LIB=$PWD/$lib python -W ignore scripts/quick_rate.py ${N_:-100000} ${WL_:-c2_gauss2d_11} $G 40 2>&1 | tail -1
  done
done
done
