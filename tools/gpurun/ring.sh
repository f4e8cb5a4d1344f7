mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_parity.py -m gpu -q -x -k "pipelined" 2>&1 | tail -3
run() {
timeout 300 python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu "$@" 2>gpurun_out/bench_tmp.err | tail -1 > gpurun_out/bench_tmp.json
python - "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/bench_tmp.json').read())
    print(' '.join(sys.argv[1:]), '| value %.1f sweeps/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel ms %.2f' % d['roofline']['kernel_ms_per_sweep'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % d['e2e']['value'], d['clocks'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/bench_tmp.err').read()[-1500:])
PY
}
run --chain-ctas 2 --panel 1984
run --chain-ctas 2 --panel 2048
timeout 300 python tools/phase_probe.py --panel 1984 --lag 1 --chain-ctas 2 --sweeps 8 2>&1 | tail -1
