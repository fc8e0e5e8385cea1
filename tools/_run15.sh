set -x
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab15_ice.log 2>&1
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 32 default > gpurun_out/r2_ab15_ice32.log 2>&1
AB_SCRIPT=ai_bench.py timeout 900 python tools/ab_variants.py run 64 corrected > gpurun_out/r2_ab15_ice_corr.log 2>&1
QB_NX=540 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_tile --launch-skip 4 -c 1 -f -o gpurun_out/r02_tile_slab python tools/quick_bench.py 64 default > gpurun_out/r2_ncu15_slab.log 2>&1
