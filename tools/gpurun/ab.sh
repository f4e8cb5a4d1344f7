for LAG in 0 1; do for P in 1024 2048 4096; do
python bench.py --steps 20 --warmup 5 --burnin 40 --no-cpu --lag $LAG --panel $P 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('lag', d['config']['lag'], 'panel', d['config']['panel'], 'value %.1f sweeps/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel ms %.2f' % d['roofline']['kernel_ms_per_sweep'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.1f' % d['e2e']['value'], 'setup %.1fs' % d['config']['setup_s'], d['clocks'])
"
done; done
