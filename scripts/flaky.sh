run() { # label, env...
  local label=$1; shift; local fails=0
  for i in 1 2 3 4 5; do
    if ! env "$@" V4H_LAUNCH_SYNC=1 timeout 100 python scripts/ds3_fwd.py 64 > /tmp/o.log 2>&1; then fails=$((fails+1)); grep "v4h\] gemm" /tmp/o.log | tail -1 | cut -c1-80; fi
  done
  echo "$label: $fails / 5 failed"
}
run base X=1
run no_tma_store V4H_GEMM_TMA_STORE=0
run stages4 V4H_GEMM_MAX_STAGES=4
run no_umma_attn V4H_DISABLE_UMMA_ATTN=1
