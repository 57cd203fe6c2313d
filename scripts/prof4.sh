CMD="python bench.py --config ds3 --steps 1 --warmup 3 --no-sampling --no-cpu-baseline --no-graph"
$CMD > gpurun_out/p4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ln_mod_bwd_vec_kernel" -s 40 -c 2 -o gpurun_out/p4_ln $CMD > gpurun_out/p4_ncu.log 2>&1
tail -n 2 gpurun_out/p4_ncu.log
