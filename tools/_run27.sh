set -x
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab27_ice.log 2>&1
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 64 corrected > gpurun_out/r2_ab27_ice_corr.log 2>&1
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 32 default > gpurun_out/r2_ab27_ice32.log 2>&1
