set -x
for s in 0 1 2; do echo "--- stagger $s"; for nx in 4320 540 1080; do QB_NX=$nx COFLUX_STAGGER=$s timeout 200 python tools/quick_bench.py 64 default; done; QB_NX=540 COFLUX_STAGGER=$s timeout 200 python tools/quick_bench.py 64 corrected; QB_NX=540 COFLUX_STAGGER=$s timeout 200 python tools/quick_bench.py 32 default; done > gpurun_out/r2_ab14.log 2>&1
COFLUX_STAGGER=1 timeout 900 python -m pytest tests/test_full_size.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest14_stagger.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ice_tile --launch-skip 2 -c 1 -f -o gpurun_out/r02_ice_tile python tools/ai_bench.py 64 default > gpurun_out/r2_ncu14_ice.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ice_ocean_bulk --launch-skip 2 -c 1 -f -o gpurun_out/r02_ice_ocean_bulk python tools/io_bench.py 64 > gpurun_out/r2_ncu14_io.log 2>&1
