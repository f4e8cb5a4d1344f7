N=${1:-2}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu "$@" > gpurun_out/m${N}_$tag.json 2> gpurun_out/m${N}_$tag.err
  python - "$tag" "$N" "$@" <<PY
import json, sys
try:
    d = json.loads(open('gpurun_out/m%s_%s.json' % (sys.argv[2], sys.argv[1])).read().strip().splitlines()[-1])
    print('N=%s' % sys.argv[2], sys.argv[1], ' '.join(sys.argv[3:]), '| value %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['state_crc'])
except Exception as e:
    print(' '.join(sys.argv[1:]), 'FAILED', e); print(open('gpurun_out/m%s_%s.err' % (sys.argv[2], sys.argv[1])).read()[-1200:])
PY
}
run c4i
JWAS_B200_LIB=$PWD/build_ab/lib_LLNI.so run c4ni
JWAS_B200_LIB=$PWD/build_ab/lib_LL8.so run c8i
JWAS_B200_LIB=$PWD/build_ab/lib_LLNI.so run c4ni_ws0 --opt ws=0
JWAS_B200_LIB=$PWD/build_ab/lib_LLNI.so run c4ni_p2048 --panel 2048 --chain-ctas 4
