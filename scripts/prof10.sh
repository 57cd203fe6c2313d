# round-2 final profile, part 3: --set full capture of the fused LayerNorm GEMM and the attention forward
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-fp32 --no-graph"
$CMD > gpurun_out/p10_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_gate_res_ln|attn_fwd" -s 48 -c 12 -o gpurun_out/p10_ln $CMD > gpurun_out/p10_ncu.log 2>&1
tail -n 2 gpurun_out/p10_ncu.log
