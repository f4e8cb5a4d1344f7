mkdir -p gpurun_out
for N in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 > gpurun_out/scale_n400k_$N.json 2> gpurun_out/scale_n400k_$N.err
done
for f in gpurun_out/scale_*.json; do echo $f; tail -1 $f | cut -c1-260; done
tail -3 gpurun_out/scale_n400k_2.err
