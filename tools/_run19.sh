set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_pytest19.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err
