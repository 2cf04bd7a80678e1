# usage: ab_envvar.sh VAR v1 v2 ... ; quick_rate.py under each value of an environment variable (development aid)
cd /root/repo
var=$1; shift
for rep in 1 2; do
for v in "$@"; do
  for G in ${GROUPS_:-2 4}; do
    echo -n "$var=$v "; env $var=$v python -W ignore scripts/quick_rate.py ${N_:-100000} ${WL_:-c2_gauss2d_11} $G 40 2>&1 | tail -1
  done
done
done
