mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_parity.py -q -x -k "warp_specialised or lag2_other or bayesa" --durations=5 2>&1 | tail -9
run() {
  tag=$1; shift
  timeout 300 python bench.py --warmup 3 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/wi_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/wi_%s.json' % sys.argv[1]).read())
    print(sys.argv[1], ' '.join(sys.argv[2:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'model %.0f act %.0f rounds %.0f' % (d['markers_in_model'], d['active_updates_per_sweep'], d['chain_rounds_per_sweep']), d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run base --steps 20 --burnin 40
run fixedpi --fixed-pi --steps 5 --burnin 20
run fixedpi_c12 --fixed-pi --steps 5 --burnin 20 --chain-ctas 12
JWAS_B200_LIB=$PWD/build_ab/lib_WIN32.so run w32_fixedpi_c12 --fixed-pi --steps 5 --burnin 20 --chain-ctas 12
JWAS_B200_LIB=$PWD/build_ab/lib_WIN128.so run w128_fixedpi_c12 --fixed-pi --steps 5 --burnin 20 --chain-ctas 12
run pi0_c12 --fixed-pi --pi0 0.0 --steps 2 --burnin 1 --chain-ctas 12
run cfg3 --config cfg3 --steps 5 --burnin 20 --chain-ctas 12
