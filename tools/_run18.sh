set -x
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench18_n8.json 2> gpurun_out/r2_bench18_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench18_n4.json 2> gpurun_out/r2_bench18_n4.err
