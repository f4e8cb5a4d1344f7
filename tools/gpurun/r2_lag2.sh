mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_parity.py -q -x -k "lag2" 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/l2_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/l2_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % d['e2e']['value'], 'model %.0f act %.0f' % (d['markers_in_model'], d['active_updates_per_sweep']), d['clocks']['sm_mhz'], d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run lag1c2 --lag 1 --chain-ctas 2
run lag2c2 --lag 2 --chain-ctas 2
run lag2c4 --lag 2 --chain-ctas 4
run lag2c6 --lag 2 --chain-ctas 6
run lag2c4p1024 --lag 2 --chain-ctas 4 --panel 1024
run lag2c8p1024 --lag 2 --chain-ctas 8 --panel 1024
run lag2c4p3072 --lag 2 --chain-ctas 4 --panel 3072
run ind --schedule independent --steps 5
run blk --schedule block --steps 3 --burnin 5
