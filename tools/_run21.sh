set -x
for nx in 4320 540; do QB_NX=$nx timeout 200 python tools/quick_bench.py 64 default; QB_NX=$nx timeout 200 python tools/quick_bench.py 64 corrected; QB_NX=$nx timeout 200 python tools/quick_bench.py 32 default; done > gpurun_out/r2_qb21.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_pytest21.log 2>&1
