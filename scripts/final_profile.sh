# End-of-round profile run (profiles/r2_final_*): the plain bench first, then its launch list under ncu, then the full captures.
cd /root/repo
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2f_plain.log 2>&1 || { echo plain bench failed; tail -3 gpurun_out/r2f_plain.log; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2f_ncu_launch.log 2>&1
wc -l gpurun_out/r2f_launches.csv
KEEP_REP=1 bash scripts/ncu_capture.sh r2f_g2 c2_gauss2d_11 mle 100000 2
bash scripts/ncu_capture.sh r2f_g4 c2_gauss2d_11 mle 100000 4
bash scripts/ncu_capture.sh r2f_c4 c4_ellip_15 mle 100000 0
bash scripts/ncu_capture.sh r2f_c5 c5_multiexp_256 mle 100000 0
