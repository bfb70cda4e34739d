mkdir -p gpurun_out
timeout 150 python scripts/soak.py 80 > gpurun_out/r02_soak_v3.txt 2>&1; echo "soak rc=$?"; tail -2 gpurun_out/r02_soak_v3.txt
for w in scan scan_p01 maxsim dense_batch; do
  case $w in scan|scan_p01) k=dense_scan;; maxsim) k=maxsim_tc5;; dense_batch) k=dense_tc5;; esac
  timeout 120 python scripts/profile_kernels.py $w > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02f_$w python scripts/profile_kernels.py $w > gpurun_out/r02f_${w}_ncu.log 2>&1
  echo "ncu $w rc=$?"
done
timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/lb.json 2>/dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_bench_launches_ncu.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/lb_ncu.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep
