mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep_parity.py -q -x -k "small_chain_units" 2>&1 | tail -3
run() {
  tag=$1; shift
  timeout 300 python bench.py --warmup 3 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/un_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/un_%s.json' % sys.argv[1]).read())
    print(sys.argv[1], ' '.join(sys.argv[2:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'model %.0f act %.0f' % (d['markers_in_model'], d['active_updates_per_sweep']), d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run base --steps 20 --burnin 40
run fixedpi_u256_c12 --fixed-pi --steps 5 --burnin 20 --chain-ctas 12 --opt unit=256
run fixedpi_u256_c24 --fixed-pi --steps 5 --burnin 20 --chain-ctas 24 --opt unit=256
run fixedpi_u256_c24_p1024 --fixed-pi --steps 5 --burnin 20 --chain-ctas 24 --opt unit=256 --panel 1024
run pi0_u256_c24 --fixed-pi --pi0 0.0 --steps 2 --burnin 1 --chain-ctas 24 --opt unit=256
run base_u256_c12 --steps 20 --burnin 40 --chain-ctas 12 --opt unit=256
