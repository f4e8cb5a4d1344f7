mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/multigpu_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -14
for C in 0 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$C bench.py --gpus 2 --steps 10 --warmup 3 --burnin 30 --chain-ctas $C > gpurun_out/scale2_c$C.json 2> gpurun_out/scale2_c$C.err
tail -1 gpurun_out/scale2_c$C.json | cut -c1-200; tail -2 gpurun_out/scale2_c$C.err | cut -c1-300
done
