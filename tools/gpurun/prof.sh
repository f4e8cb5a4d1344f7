# round-1 final measurements (one B200): full GPU test suite, smoke, bench (both arms), launch list, one full ncu capture
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --burnin 40 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
tail -c 2800 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r1_reference.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --burnin 40 --fixed-pi --no-cpu > gpurun_out/bench_r1_fixedpi.json 2>/dev/null
tail -c 300 gpurun_out/bench_r1_fixedpi.json
python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu --gather --panel 3968 --chain-ctas 4 > gpurun_out/bench_r1_gather3968.json 2>/dev/null
tail -c 300 gpurun_out/bench_r1_gather3968.json
python tools/method_probe.py --method R --sweeps 10 2>&1 | tail -2
python tools/method_probe.py --method MT --n 20000 --p 500000 --sweeps 10 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --burnin 40 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 44 -c 1 -f -o gpurun_out/prof_fused_r1 python bench.py --steps 1 --warmup 3 --burnin 40 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -c 200 gpurun_out/ncu_full.log
