set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_v2_gpu.py -m gpu -x -q -k "sea_ice or config5 or averaged" 2>&1 | tail -8 > gpurun_out/r2_pytest9.log
timeout 300 python tools/io_bench.py 64 > gpurun_out/r2_io9.log 2>&1
COFLUX_IO_BULK=0 timeout 300 python tools/io_bench.py 64 >> gpurun_out/r2_io9.log 2>&1
timeout 300 python tools/io_bench.py 32 >> gpurun_out/r2_io9.log 2>&1
COFLUX_IO_BULK=0 timeout 300 python tools/io_bench.py 32 >> gpurun_out/r2_io9.log 2>&1
timeout 900 python -m pytest tests/test_full_size.py -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/r2_pytest9.log
