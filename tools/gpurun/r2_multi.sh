# round 2: N-GPU correctness (bit-identity with one GPU) and a short scaling probe; N = first argument
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\[W\|warn\|OMP_NUM\|\*\*\*\*" | tail -16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/weak_n$N.json 2> gpurun_out/weak_n$N.err
tail -1 gpurun_out/weak_n$N.json | cut -c1-400; grep -v "^W\|^\[W\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/weak_n$N.err | tail -5 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --strong --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err
tail -1 gpurun_out/strong_n$N.json | cut -c1-400; grep -v "^W\|^\[W\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/strong_n$N.err | tail -5 | cut -c1-300
