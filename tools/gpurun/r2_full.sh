# round-2 measurements on one B200: full GPU test suite, smoke, bench (both arms), launch list, full ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 600 gpurun_out/bench_r2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --burnin 40 --no-cpu --no-extras > gpurun_out/ncu_bench_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 44 -c 1 -f -o gpurun_out/prof_fused_r2 python bench.py --steps 1 --warmup 3 --burnin 40 --no-cpu --no-extras > gpurun_out/ncu_full_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jw_k_stream -s 8 -c 1 -f -o gpurun_out/prof_stream_r2 python bench.py --steps 1 --warmup 3 --burnin 8 --schedule independent --no-cpu --no-extras > gpurun_out/ncu_full_stream_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
