set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline"
$CMD > gpurun_out/p1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_.*umma -s 36 -c 3 -o gpurun_out/p1_attn $CMD > gpurun_out/p1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 400 -c 40 -o gpurun_out/p1_gemm $CMD > gpurun_out/p1_ncu2.log 2>&1
tail -n 3 gpurun_out/p1_ncu1.log gpurun_out/p1_ncu2.log
grep -o '"value": [0-9.]*' gpurun_out/p1_plain.log | head -2
