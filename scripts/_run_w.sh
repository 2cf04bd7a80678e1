cd /root/repo
for rep in 1 2; do
for v in base w2 w1 w6; do
  if [ $v = base ]; then lib=dphtools_b200/csrc/libdphlm.so; pc=8; fi
  if [ $v = w2 ]; then lib=dphtools_b200/csrc/variants/w2.so; pc=16; fi
  if [ $v = w1 ]; then lib=dphtools_b200/csrc/variants/w1.so; pc=32; fi
  if [ $v = w6 ]; then lib=dphtools_b200/csrc/variants/w6.so; pc=6; fi
  for G in 2 4; do
    DPHLM_LIB=$PWD/$lib DPHLM_POLLER_CTAS=$pc python -W ignore scripts/quick_rate.py 100000 c2_gauss2d_11 $G 40 2>&1 | tail -1
  done
done
done
