# usage: ab_envvar_wl.sh VAR "v1 v2" "workload:lanes ..." ; quick_rate.py per workload under each value of VAR (development aid)
cd /root/repo
var=$1; vals=$2; shift; shift
for rep in 1 2; do
for wg in "$@"; do
  wl=${wg%%:*}; G=${wg##*:}
  for v in $vals; do
    if [ "$v" = unset ]; then echo -n "$var unset "; env -u $var python -W ignore scripts/quick_rate.py ${N_:-100000} $wl $G 30 2>&1 | tail -1
    else echo -n "$var=$v "; env $var=$v python -W ignore scripts/quick_rate.py ${N_:-100000} $wl $G 30 2>&1 | tail -1; fi
  done
done
done
