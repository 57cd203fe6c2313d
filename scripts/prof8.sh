# round-2 final profile, part 1: per-launch durations of the ds2 training step (eager launches: ncu cannot see inside a
# replayed graph).  Plain run first, then the same command under ncu.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-sampling --no-cpu-baseline --no-fp32 --no-graph"
$CMD > gpurun_out/p8_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 600 --csv --log-file gpurun_out/p8_launches.csv $CMD > gpurun_out/p8_ncu1.log 2>&1
tail -n 2 gpurun_out/p8_ncu1.log
