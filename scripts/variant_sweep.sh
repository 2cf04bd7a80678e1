#!/bin/bash
# run the profile driver against every library variant under dphtools_b200/csrc/variants (tuning aid)
for lib in dphtools_b200/csrc/libdphlm.so dphtools_b200/csrc/variants/*.so; do
  echo "== $lib"
  DPHLM_LIB=$PWD/$lib timeout 300 python -W ignore scripts/profile_run.py "$@" 2>&1 | tail -n 4
done
