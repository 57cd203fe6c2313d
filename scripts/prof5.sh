set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-graph"
$CMD > gpurun_out/p5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 320 --csv --log-file gpurun_out/p5_launches.csv $CMD > gpurun_out/p5_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_umma|attn_" -s 320 -c 24 -o gpurun_out/p5_top $CMD > gpurun_out/p5_ncu2.log 2>&1
tail -n 2 gpurun_out/p5_ncu1.log gpurun_out/p5_ncu2.log
