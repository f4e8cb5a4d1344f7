# A/B of engine-1 options at cfg2 on one box (every run prints one summary line)
mkdir -p gpurun_out
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
run --chain-ctas 0 --panel 2048
run --chain-ctas 2 --panel 2048
run --chain-ctas 2 --panel 1984 --gather
run --chain-ctas 4 --panel 3968 --gather
# a second build of the library (another commit, or JWAS_B200_BUILD_FLAGS=-DJW_TIMERS) is selected with JWAS_B200_LIB=<path>
