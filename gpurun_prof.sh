set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --burnin 40 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
tail -c 3000 gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --burnin 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 8 -c 1 -o gpurun_out/prof_fused_r1 python bench.py --steps 1 --warmup 3 --burnin 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
