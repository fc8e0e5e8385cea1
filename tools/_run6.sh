set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest6.log
COFLUX_LIB=climaocean.jl_b200/lib/variants/ice_coare.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_v2_gpu.py -m gpu -q -k "sea_ice" 2>&1 | tail -25 > gpurun_out/r2_pytest6_icecoare.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
