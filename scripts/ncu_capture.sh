#!/bin/bash
# ncu --set full capture of the four stage kernels of one call (the last of three), per the profiling recipe:
# the plain command first, then the same command under ncu.  usage: ncu_capture.sh TAG WORKLOAD METHOD N LANES
# Output: gpurun_out/TAG.ncu-rep (summarise here with scripts/ncu_summary.py) and gpurun_out/TAG_plain.log
set -u
tag=$1; wl=$2; method=$3; n=$4; lanes=$5
mkdir -p gpurun_out
python -W ignore scripts/profile_run.py $wl $method $n --lanes $lanes --reps 3 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -1 gpurun_out/${tag}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dphlm_stage_kernel --launch-skip 8 --launch-count 4 \
  -f -o gpurun_out/$tag python -W ignore scripts/profile_run.py $wl $method $n --lanes $lanes --reps 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
# the report itself is 15-20 MB (gpurun_out is limited to 64 MiB): summarise on the box, keep the report only if KEEP_REP is set
python scripts/ncu_summary.py gpurun_out/$tag.ncu-rep --blocks > gpurun_out/$tag.txt 2>&1
[ -n "${KEEP_REP:-}" ] || rm -f gpurun_out/$tag.ncu-rep
