set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --burnin 40 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
tail -c 2500 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r1_reference.json 2>/dev/null
cat gpurun_out/bench_r1_reference.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --burnin 30 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -c 400 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 34 -c 1 -o gpurun_out/prof_fused_r1 python bench.py --steps 1 --warmup 3 --burnin 30 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -c 300 gpurun_out/ncu_full.log
