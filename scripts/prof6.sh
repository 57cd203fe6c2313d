# round-1 (d) profile: bench line, then ONE ncu --set full capture of the LayerNorm / attention kernels
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-graph"
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/p6_bench.json 2> gpurun_out/p6_bench.err
python scripts/bench_brief.py gpurun_out/p6_bench.json 6
$CMD > gpurun_out/p6_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ln_mod|attn_" -s 150 -c 16 -o gpurun_out/p6_ln_attn $CMD > gpurun_out/p6_ncu.log 2>&1
tail -n 2 gpurun_out/p6_ncu.log
