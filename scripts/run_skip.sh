for sk in 0 1 2 3 4 7; do echo "skip=$sk"; V4H_GEMM_DBG_SKIP=$sk GEMM_M=34560 timeout 100 python scripts/gemm_bench.py 2>&1 | sed -n 2,3p | cut -c1-330; done
