# round-2 final profile, part 2: ONE --set full capture of the dominant kernels of the same command
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-fp32 --no-graph"
$CMD > gpurun_out/p9_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_umma|gemm_gate_res_ln|attn_|ln_mod" -s 300 -c 40 -o gpurun_out/p9_top $CMD > gpurun_out/p9_ncu.log 2>&1
tail -n 2 gpurun_out/p9_ncu.log
