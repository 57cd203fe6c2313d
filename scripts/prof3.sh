set -x
CMD="python bench.py --steps 1 --warmup 3 --no-sampling --no-cpu-baseline --no-graph"
$CMD > gpurun_out/p3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ln_mod_bwd_vec_kernel|ln_mod_fwd_vec|adamw_kernel" -s 120 -c 6 -o gpurun_out/p3_ln $CMD > gpurun_out/p3_ncu.log 2>&1
tail -n 2 gpurun_out/p3_ncu.log
