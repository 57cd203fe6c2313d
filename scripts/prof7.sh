# round-1 final profile of the ds2 training step (eager launches: ncu cannot see inside a replayed graph):
# plain run first, then the per-launch duration list, then ONE --set full capture of the dominant kernels
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-graph"
$CMD > gpurun_out/p7_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 330 --csv --log-file gpurun_out/p7_launches.csv $CMD > gpurun_out/p7_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_umma|attn_|ln_mod" -s 330 -c 30 -o gpurun_out/p7_top $CMD > gpurun_out/p7_ncu2.log 2>&1
tail -n 2 gpurun_out/p7_ncu1.log gpurun_out/p7_ncu2.log
