set -x
timeout 400 python tools/ab_variants.py run 64 default corrected > gpurun_out/r2_ab8.log 2>&1
QB_NX=540 timeout 400 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab8_slab.log 2>&1
COFLUX_BALANCE=0 QB_NX=540 timeout 200 python tools/quick_bench.py 64 default > gpurun_out/r2_ab8_nobalance.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest8.log
