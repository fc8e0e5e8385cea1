set -x
AB_SCRIPT=io_bench.py COFLUX_IO_BULK=1 timeout 600 python tools/ab_variants.py run 64 > gpurun_out/r2_io10.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_v2_gpu.py -m gpu -x -q -k "sea_ice or config5" 2>&1 | tail -8 > gpurun_out/r2_pytest10.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flux_tile -s 1 -c 1 -o gpurun_out/r2_tile_c python tools/profile_step.py 64 default 3 twelfth > gpurun_out/r2_ncu_d.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_update_state_config1 and default and 64" > gpurun_out/r02_racecheck_tile.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_atmosphere_sea_ice_fluxes and default-64" > gpurun_out/r02_racecheck_ice_tile.log 2>&1
