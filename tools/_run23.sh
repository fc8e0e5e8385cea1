set -x
for cfg in default corrected; do for bits in 64 32; do
COFLUX_ICE_QUEUE=0 timeout 300 python tools/ice_dump.py /tmp/ice_old_${cfg}_$bits.npz $bits $cfg
timeout 300 python tools/ice_dump.py /tmp/ice_new_${cfg}_$bits.npz $bits $cfg
python tools/ice_cmp.py /tmp/ice_old_${cfg}_$bits.npz /tmp/ice_new_${cfg}_$bits.npz
done; done > gpurun_out/r2_ice23_cmp.log 2>&1
for cfg in default corrected ncar; do timeout 300 python tools/ai_bench.py 64 $cfg; COFLUX_ICE_QUEUE=0 timeout 300 python tools/ai_bench.py 64 $cfg; done > gpurun_out/r2_ice23_bench.log 2>&1
timeout 300 python tools/ai_bench.py 32 default >> gpurun_out/r2_ice23_bench.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q -k "ice or coupled or config5 or averaged" 2>&1 | tail -6 ) > gpurun_out/r2_pytest23.log 2>&1
