set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest3.log
timeout 120 python tools/quick_bench.py 64 default corrected > gpurun_out/r2_qb3.log 2>&1
COFLUX_KERNEL=tile timeout 120 python tools/quick_bench.py 64 default corrected >> gpurun_out/r2_qb3.log 2>&1
timeout 120 python tools/quick_bench.py 32 default corrected >> gpurun_out/r2_qb3.log 2>&1
timeout 300 python -m pytest tests/test_full_size.py -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/r2_pytest3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flux_stream -s 1 -c 1 -o gpurun_out/r2_stream_a python tools/profile_step.py 64 default 3 twelfth > gpurun_out/r2_ncu_c.log 2>&1
