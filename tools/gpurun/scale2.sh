mkdir -p gpurun_out
N=2
for C in 0 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$C bench.py --gpus $N --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 --chain-ctas $C > gpurun_out/scale_n400k_${N}_c$C.json 2> gpurun_out/scale_n400k_${N}_c$C.err
tail -1 gpurun_out/scale_n400k_${N}_c$C.json | cut -c1-200
done
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 --no-cpu --chain-ctas 0 2>/dev/null | tail -1 | cut -c1-200
