# small launches: the deferred stage (two kernels) against the single raster kernel, 1 / 4 / 16 views per step
cd /root/repo
for defer in 1 0; do for w in mesh1080 views1080_tex; do for v in 1 4 16; do
DTR_B200_DEFER=$defer python bench.py --workload $w --views $v --steps 200 --warmup 10 --e2e-steps 2 --no-cpu-baseline --no-others 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('defer=$defer $w views $v', 'ms/step', round(d['ms_per_step'],4), 'raster', round(r['stage_ms_per_step']['raster'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'parity', d['parity_checked'])"
done; done; done
