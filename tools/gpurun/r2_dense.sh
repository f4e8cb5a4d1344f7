mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep_parity.py -q -x -k "lag2 or pipelined or bayesa or external or host_array" 2>&1 | tail -3
run() {
  tag=$1; shift
  timeout 300 python bench.py --warmup 3 --no-cpu --no-extras "$@" 2>gpurun_out/ab_tmp.err | tail -1 > gpurun_out/de_$tag.json
  python - "$tag" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/de_%s.json' % sys.argv[1]).read())
    print(' '.join(sys.argv[1:]), '| value %.1f' % d['value'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'model %.0f act %.0f rounds %.0f' % (d['markers_in_model'], d['active_updates_per_sweep'], d['chain_rounds_per_sweep']), d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/ab_tmp.err').read()[-800:])
PY
}
run base --steps 20 --burnin 40
run fixedpi_c4 --fixed-pi --steps 5 --burnin 20
run fixedpi_c8 --fixed-pi --steps 5 --burnin 20 --chain-ctas 8
run fixedpi_c8_l1 --fixed-pi --steps 5 --burnin 20 --chain-ctas 8 --lag 1
run pi0_c8 --fixed-pi --pi0 0.0 --steps 2 --burnin 1 --chain-ctas 8
export JWAS_B200_LIB=$PWD/build_ab/lib_RB8.so
run rb8_base --steps 20 --burnin 40
run rb8_fixedpi_c8 --fixed-pi --steps 5 --burnin 20 --chain-ctas 8
run rb8_pi0_c8 --fixed-pi --pi0 0.0 --steps 2 --burnin 1 --chain-ctas 8
