N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus $N --steps 5 --warmup 3 --burnin 5 --no-extras > gpurun_out/cpuleg_n$N.json 2> gpurun_out/cpuleg_n$N.err
tail -1 gpurun_out/cpuleg_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['cpu_baseline']
print('cpu value', c['value'], 'cores', c['cores'], 'affinity', c.get('affinity_cores'), c['sample'][-60:], 'one_thread', c.get('one_thread'))"
tail -3 gpurun_out/cpuleg_n$N.err | cut -c1-200
