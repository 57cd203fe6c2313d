set -x
CMD="python bench.py --steps 3 --warmup 3 --no-sampling --no-cpu-baseline"
$CMD > gpurun_out/p0_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/p0_launches.csv $CMD > gpurun_out/p0_ncu1.log 2>&1
$CMD > gpurun_out/p0_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 300 -c 3 -o gpurun_out/p0_gemm $CMD > gpurun_out/p0_ncu2.log 2>&1
tail -3 gpurun_out/p0_ncu1.log gpurun_out/p0_ncu2.log
