# the driver's SCALE command at N GPUs (with the nested results), plus the reference arm under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
tail -1 gpurun_out/scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.1f problem %.1f ms/step %.2f kernel %.3f frac %.3f e2e %.1f scaling %s crc %s' % (d['n_gpus'], d['value'], d['problem_sweeps_per_s'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['scaling'], d['state_crc']))
for k in ('strong',):
    v=d.get(k); print(k, {kk: v.get(kk) for kk in ('sweeps_per_s','ms_per_step','state_crc','error')} if v else None)
for k,v in d.get('configs',{}).items(): print(k, {kk: v.get(kk) for kk in ('sweeps_per_s','ms_per_step','markers_in_model','state_crc','setup_s','error')})
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --impl reference --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-330
