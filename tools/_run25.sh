set -x
for cfg in default corrected; do timeout 200 python tools/quick_bench.py 32 $cfg; done > gpurun_out/r2_qb25.log 2>&1
timeout 200 python tools/quick_bench.py 64 ncar >> gpurun_out/r2_qb25.log 2>&1
timeout 200 python tools/quick_bench.py 32 ncar >> gpurun_out/r2_qb25.log 2>&1
timeout 200 python tools/quick_bench.py 64 default >> gpurun_out/r2_qb25.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_pytest25.log 2>&1
timeout 600 python tests/diag/parity_report.py > gpurun_out/r02_parity_report_final.log 2>&1
