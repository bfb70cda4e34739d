# re-capture after the last changes to dense_scan.cu (merge floor, tournament limit) and config 3 at its full size
mkdir -p gpurun_out
for w in scan scan_p01 dense_batch_10m; do
  case $w in scan|scan_p01) k=dense_scan;; *) k=dense_tc5;; esac
  timeout 200 python scripts/profile_kernels.py $w 2 > /dev/null 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02f_$w python scripts/profile_kernels.py $w 2 > gpurun_out/r02f_${w}_ncu.log 2>&1
  echo "ncu $w rc=$?"
done
timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/lb.json 2>/dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_bench_launches_ncu.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/lb_ncu.log 2>&1
echo "launch list rc=$?"
