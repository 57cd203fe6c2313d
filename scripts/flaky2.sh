run() { local label=$1; shift; local fails=0
  for i in 1 2 3 4; do
    if ! env "$@" REPS=1 timeout 100 python scripts/ds3_sample.py > /tmp/o.log 2>&1; then fails=$((fails+1)); grep "v4h\] gemm\|^rep" /tmp/o.log | tail -2 | cut -c1-90; fi
  done
  echo "$label: $fails / 4 failed"
}
run sync V4H_LAUNCH_SYNC=1
run nosync X=1
run sync_stages4 V4H_LAUNCH_SYNC=1 V4H_GEMM_MAX_STAGES=4
run sync_notma V4H_LAUNCH_SYNC=1 V4H_GEMM_TMA_STORE=0
