cd /root/repo
for conn in 8 32; do
for ch in auto 50000 100000; do
  if [ $ch = auto ]; then unset DPHLM_CHUNK; else export DPHLM_CHUNK=$ch; fi
  echo "CUDA_DEVICE_MAX_CONNECTIONS=$conn"; CUDA_DEVICE_MAX_CONNECTIONS=$conn python -W ignore scripts/e2e_rate.py 100000 10 2>&1 | tail -1
done
done
