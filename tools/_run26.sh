set -x
for cfg in default corrected ncar; do timeout 300 python tools/ai_bench.py 64 $cfg; done > gpurun_out/r2_ice26_bench.log 2>&1
timeout 300 python tools/ai_bench.py 32 default >> gpurun_out/r2_ice26_bench.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q -k "ice or coupled or config5 or averaged" 2>&1 | tail -6 ) > gpurun_out/r2_pytest26.log 2>&1
