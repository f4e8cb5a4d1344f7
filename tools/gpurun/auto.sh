timeout 600 python -m pytest tests/test_gpu_sweep_parity.py -m gpu -q -x -k "pipelined_chain and (60013 or 500-2000 or 300-9000)" 2>&1 | tail -2
timeout 300 python bench.py --nobs 400000 --nmarkers 61440 --steps 10 --warmup 3 --burnin 25 --no-cpu 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --burnin 40 --no-cpu 2>/dev/null | tail -1 | cut -c1-200
