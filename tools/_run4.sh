set -x
timeout 600 python -m pytest tests/test_v2_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest4.log
timeout 400 python tools/ab_variants.py run 64 default corrected > gpurun_out/r2_ab4_f64.log 2>&1
timeout 300 python tools/ab_variants.py run 32 default corrected > gpurun_out/r2_ab4_f32.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest4_all.log
