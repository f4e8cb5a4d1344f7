mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --warmup 3 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/po_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/po_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'model %.0f act %.0f' % (d['markers_in_model'], d['active_updates_per_sweep']), d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
for ss in 0 300 1000; do for sc in 0 100; do
run base_s${ss}_c${sc} --steps 20 --burnin 40 --opt poll_ns_stream=$ss --opt poll_ns_chain=$sc
run fixedpi_s${ss}_c${sc} --fixed-pi --steps 5 --burnin 20 --chain-ctas 8 --opt poll_ns_stream=$ss --opt poll_ns_chain=$sc
done; done
run pi0_s1000_c100 --fixed-pi --pi0 0.0 --steps 2 --burnin 1 --chain-ctas 8 --opt poll_ns_stream=1000 --opt poll_ns_chain=100
