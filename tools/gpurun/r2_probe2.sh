export JWAS_B200_LIB=$PWD/build_ab/lib_TIMERS.so
PROBE_PI=0.95 PROBE_LAG=2 PROBE_CC=4 python tools/multi_phase_probe.py 2>&1 | tail -2
PROBE_PI=0.95 PROBE_LAG=1 PROBE_CC=2 python tools/multi_phase_probe.py 2>&1 | tail -2
PROBE_PI=0.999 PROBE_LAG=2 PROBE_CC=4 python tools/multi_phase_probe.py 2>&1 | tail -2
