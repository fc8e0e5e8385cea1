set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest5.log
COFLUX_LIB=climaocean.jl_b200/lib/variants/ice_coare.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_v2_gpu.py -m gpu -q -k "sea_ice" 2>&1 | tail -25 > gpurun_out/r2_pytest5_icecoare.log
timeout 600 python tests/diag/parity_report.py > gpurun_out/r02_parity_report.log 2>&1
timeout 200 python tools/quick_bench.py 64 default corrected ncar > gpurun_out/r2_qb5.log 2>&1
timeout 200 python tools/quick_bench.py 32 default corrected ncar >> gpurun_out/r2_qb5.log 2>&1
