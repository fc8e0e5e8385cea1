set -x
( time timeout 1500 python -m pytest tests/test_full_size.py tests/test_normalize_salinity.py -m gpu -x -q -s 2>&1 | tail -25 ) > gpurun_out/r2_pytest13.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_ref13.json 2> gpurun_out/r2_ref13.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu13_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ice_ocean_bulk|normalize_salinity|net_sea_ice" -c 6 -f -o gpurun_out/r02_aux python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu13_aux.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_tile --launch-skip 4 -c 1 -f -o gpurun_out/r02_tile_final python tools/quick_bench.py 64 default > gpurun_out/r2_ncu13_tile.log 2>&1
ls -la gpurun_out
