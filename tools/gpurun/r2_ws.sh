mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --warmup 3 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/ws_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/ws_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % (d['e2e']['value'] if d['e2e'] else 0), 'model %.0f act %.0f' % (d['markers_in_model'], d['active_updates_per_sweep']), d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run ws_nopf --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0
run ws_nopf_p3072 --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0 --panel 3072
run ws_nopf_p4096_c6 --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0 --panel 4096 --chain-ctas 6
run ws_nopf_c6 --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0 --chain-ctas 6
run ws_nopf_lag1 --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0 --lag 1 --chain-ctas 2
run base_nopf --steps 20 --burnin 40 --opt l2_prefetch=0
for nb in 2 3 6; do JWAS_B200_LIB=$PWD/build_ab/lib_NB$nb.so run ws_nopf_nb$nb --steps 20 --burnin 40 --opt ws=1 --opt l2_prefetch=0; done
