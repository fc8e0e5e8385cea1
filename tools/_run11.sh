set -x
for v in f32_libm f32_fast; do echo "--- $v"; COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/quick_bench.py 32 default corrected; done > gpurun_out/r2_ab11_f32.log 2>&1
COFLUX_LIB=climaocean.jl_b200/lib/variants/f32_fast.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_v2_gpu.py -m gpu -q -k "32" 2>&1 | tail -15 > gpurun_out/r2_pytest11_f32fast.log
COFLUX_LIB=climaocean.jl_b200/lib/variants/f32_fast.so timeout 600 python tests/diag/parity_report.py 2>&1 | grep -A25 "f32 default" | cut -c1-330 > gpurun_out/r2_parity11_f32fast.log
for v in iob_256_8_3 iob_256_8_2 iob_256_4_6 iob_288_8_3 iob_288_8_2 iob_288_5_5 iob_384_8_2; do echo "--- $v"; COFLUX_IO_BULK=1 COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/io_bench.py 64; COFLUX_IO_BULK=1 COFLUX_LIB=climaocean.jl_b200/lib/variants/$v.so timeout 200 python tools/io_bench.py 32; done > gpurun_out/r2_io11.log 2>&1
