set -x
for v in base_park carry2_320x2_896 carry2_320x2_960_k5 carry2_384x2_896_r80; do echo "--- $v"; COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/quick_bench.py 64 default; QB_NX=540 COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/quick_bench.py 64 default; done > gpurun_out/r2_ab12.log 2>&1
for v in base_park io32_384 io32_512 io32_576; do echo "--- $v"; COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/io_bench.py 32; done > gpurun_out/r2_io12.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest12.log
