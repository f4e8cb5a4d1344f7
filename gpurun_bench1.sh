set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --n 50000 --p 60000 --steps 5 --warmup 3 --burnin 10 --cpu-markers 1000 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --burnin 30 2>&1 | tail -5
