# final re-capture of dense_tc5_kernel (config 3 at full size, and the 2M-row shape) after the last changes to dense_tc5.cu
mkdir -p gpurun_out
for w in dense_batch dense_batch_10m; do
  timeout 200 python scripts/profile_kernels.py $w 2 > /dev/null 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_tc5 -s 1 -c 1 -f -o gpurun_out/r02f_$w python scripts/profile_kernels.py $w 2 > gpurun_out/r02f_${w}_ncu.log 2>&1
  echo "ncu $w rc=$?"
done
