set -x
timeout 900 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab17.log 2>&1
QB_NX=540 timeout 900 python tools/ab_variants.py run 64 default > gpurun_out/r2_ab17_slab.log 2>&1
