mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/ro_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/ro_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % (d['e2e']['value'] if d['e2e'] else 0), 'model %.0f act %.0f' % (d['markers_in_model'], d['active_updates_per_sweep']), d['clocks']['sm_mhz'], d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run lag1c2 --lag 1 --chain-ctas 2
run lag2c4 --lag 2 --chain-ctas 4
run lag2c3 --lag 2 --chain-ctas 3
run lag2c4p3072 --lag 2 --chain-ctas 4 --panel 3072
run lag2c4p1536 --lag 2 --chain-ctas 4 --panel 1536
run lag1c2g --lag 1 --chain-ctas 2 --panel 1984 --opt gather=1
run lag2c4g --lag 2 --chain-ctas 4 --panel 1984 --opt gather=1
run lag2c6gp3968 --lag 2 --chain-ctas 6 --panel 3968 --opt gather=1
run fixedpi_l2c4 --lag 2 --chain-ctas 4 --fixed-pi --steps 5
run fixedpi_l2c8 --lag 2 --chain-ctas 8 --fixed-pi --steps 5
