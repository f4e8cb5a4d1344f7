export JWAS_B200_LIB=$PWD/build_ab/lib_TIMERS.so
python tools/multi_phase_probe.py 2>&1 | tail -3
PROBE_N=25000 python tools/multi_phase_probe.py 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tools/multi_phase_probe.py 2>&1 | grep "^N=" | tail -3
PROBE_N=100000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 tools/multi_phase_probe.py 2>&1 | grep "^N=" | tail -3
