set -x
timeout 400 python tools/ab_variants.py run 64 default corrected > gpurun_out/r2_ab7.log 2>&1
QB_NX=540 timeout 400 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab7_slab.log 2>&1
COFLUX_TAPER=0 QB_NX=540 timeout 200 python tools/quick_bench.py 64 default > gpurun_out/r2_ab7_notaper.log 2>&1
COFLUX_TAPER=0 timeout 200 python tools/quick_bench.py 64 default >> gpurun_out/r2_ab7_notaper.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest7.log
