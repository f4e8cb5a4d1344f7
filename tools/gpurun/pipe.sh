# pipelined chain: parity tests, then A/B of chain_ctas at cfg2, then the phase probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_parity.py -m gpu -q -x -k "pipelined or mega or lagged" 2>&1 | tail -5
for C in 0 2 4 8; do
timeout 300 python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu --chain-ctas $C 2>gpurun_out/bench_c$C.err | tail -1 > gpurun_out/bench_c$C.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c$C.json').read())
    print('chain_ctas', $C, 'value %.1f sweeps/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel ms %.2f' % d['roofline']['kernel_ms_per_sweep'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % d['e2e']['value'], 'rounds %.0f' % d['config']['chain_rounds_per_sweep'], d['clocks'])
except Exception as e:
    print('chain_ctas', $C, 'FAILED', e); print(open('gpurun_out/bench_c$C.err').read()[-1500:])
PY
done
timeout 300 python tools/phase_probe.py --panel 2048 --lag 1 --chain-ctas 4 --sweeps 8 2>&1 | tail -4
timeout 300 python tools/phase_probe.py --panel 2048 --lag 1 --chain-ctas 0 --sweeps 8 2>&1 | tail -2
