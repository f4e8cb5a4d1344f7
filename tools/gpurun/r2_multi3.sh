N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/multigpu_check.py 2>&1 | grep "rank 0/\|MULTIGPU" | tail -16
run() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu "$@" > gpurun_out/m${N}_$tag.json 2> gpurun_out/m${N}_$tag.err
  python - "$tag" "$N" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/m%s_%s.json' % (sys.argv[2], sys.argv[1])).read().strip().splitlines()[-1])
    print('N=%s' % sys.argv[2], ' '.join(sys.argv[3:]), '| value %.1f' % d['value'], 'problem %.1f' % d['problem_sweeps_per_s'], 'ms/step %.2f' % d['ms_per_step'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % (d['e2e']['value'] if d['e2e'] else 0), d['state_crc'], 'setup %.1f' % d['setup_s'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/m%s_%s.err' % (sys.argv[2], sys.argv[1])).read()[-1200:])
PY
}
run weak
run weak_c8 --chain-ctas 8
run strong --strong
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 value %.1f ms/step %.2f kernel %.3f frac %.3f crc %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['state_crc']))"
