set -x
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest2.log
python tools/ab_variants.py run 64 default corrected > gpurun_out/r2_ab2_f64.log 2>&1
python tools/ab_variants.py run 32 default corrected > gpurun_out/r2_ab2_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flux_tile -s 1 -c 1 -o gpurun_out/r2_tile_b python tools/profile_step.py 64 default 3 twelfth > gpurun_out/r2_ncu_b.log 2>&1
