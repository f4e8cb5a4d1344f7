# row-sharded scaling at 400,000 x 61,440 (stream-dominated) with the pipelined chain; N = number of visible GPUs
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 --no-cpu > gpurun_out/scale_n400k_1.json 2> gpurun_out/scale_n400k_1.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 > gpurun_out/scale_n400k_$N.json 2> gpurun_out/scale_n400k_$N.err
fi
tail -1 gpurun_out/scale_n400k_$N.json | cut -c1-220; tail -2 gpurun_out/scale_n400k_$N.err | cut -c1-300
