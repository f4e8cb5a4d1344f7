# round 2, call 1: A/B of the three JW_NEXT_* switches (separate builds under build_ab/) at cfg2, plus
# dense-regime probes (fixed pi, more chain CTAs) and panel 4096
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/ab_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/ab_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_sweep'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % d['e2e']['value'], 'model %.0f act %.0f' % (d['config']['markers_in_model'], d['config']['active_updates_per_sweep']), d['clocks']['sm_mhz'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run base
for v in META L1PF RED ALL; do JWAS_B200_LIB=$PWD/build_ab/lib_$v.so run $v; done
run base2
run p4096 --panel 4096
run p1024 --panel 1024
run c4 --chain-ctas 4
run fixedpi_c2 --fixed-pi --steps 5
run fixedpi_c8 --fixed-pi --steps 5 --chain-ctas 8
run fixedpi_c16 --fixed-pi --steps 5 --chain-ctas 16
run fixedpi_c8_p1024 --fixed-pi --steps 5 --chain-ctas 8 --panel 1024
