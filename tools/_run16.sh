set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest16_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench16_n2.json 2> gpurun_out/r2_bench16_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_ref16_n2.json 2> gpurun_out/r2_ref16_n2.err
