mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 44 -c 1 -o gpurun_out/prof_pipe_noring -f python bench.py --steps 1 --warmup 3 --burnin 40 --no-cpu --chain-ctas 2 --no-ring > gpurun_out/ncu_a.log 2>&1
tail -c 300 gpurun_out/ncu_a.log
ncu --set full --clock-control none --import-source on -k regex:jw_k_fused -s 44 -c 1 -o gpurun_out/prof_pipe_ring -f python bench.py --steps 1 --warmup 3 --burnin 40 --no-cpu --chain-ctas 2 > gpurun_out/ncu_b.log 2>&1
tail -c 300 gpurun_out/ncu_b.log
ls -la gpurun_out/
