set -x
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or ragged or mixed or host_buffer" > gpurun_out/r02_memcheck_tile.log 2>&1; echo "rc $?" >> gpurun_out/r02_memcheck_tile.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_normalize_salinity.py tests/test_v2_gpu.py -m gpu -x -q -k "cuda_normalization or net_sea_ice or time_averaged or land_freshwater" > gpurun_out/r02_memcheck_v2.log 2>&1; echo "rc $?" >> gpurun_out/r02_memcheck_v2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 and 64" > gpurun_out/r02_racecheck_flux_tile_final.log 2>&1; echo "rc $?" >> gpurun_out/r02_racecheck_flux_tile_final.log
tail -5 gpurun_out/r02_memcheck_tile.log gpurun_out/r02_memcheck_v2.log gpurun_out/r02_racecheck_flux_tile_final.log
