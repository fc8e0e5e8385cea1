set -x
timeout 900 python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_ref24.json 2> gpurun_out/r2_ref24.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu24_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_tile --launch-skip 4 -c 1 -f -o gpurun_out/r02_tile_final python tools/quick_bench.py 64 default > gpurun_out/r2_ncu24_tile.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_tile --launch-skip 4 -c 1 -f -o gpurun_out/r02_tile_final_f32 python tools/quick_bench.py 32 default > gpurun_out/r2_ncu24_tile32.log 2>&1
