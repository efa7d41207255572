cd /root/repo
for lib in /root/repo/dtrenderer_b200/libdtr_b200.so "$@"; do
for v in 64 512; do
DTR_B200_LIB=$lib python bench.py --views $v --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-others 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib'.split('/')[-1], 'views $v', 'ms/step', round(d['ms_per_step'],4), 'raster', round(r['stage_ms_per_step']['raster'],4), 'iso', round(r['stage_ms_isolated']['raster'],4), 'frac', round(r['frac'],4), 'parity', d['parity_checked'])"
done; done
