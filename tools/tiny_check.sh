cd /root/repo
timeout 300 python -X faulthandler -m pytest tests -m gpu -q 2>&1 | tail -3
for w in mesh1080 views1080_tex; do for v in 1 4 16; do
python bench.py --workload $w --views $v --steps 200 --warmup 10 --e2e-steps 2 --no-cpu-baseline --no-others 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$w views $v', 'ms/step', round(d['ms_per_step'],4), 'raster', round(r['stage_ms_per_step']['raster'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'parity', d['parity_checked'])"
done; done
python tools/band_probe.py 8 4 1
python bench.py --workload views1080_tex --views 64 --steps 50 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-others 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('views1080_tex 64', 'ms/step', round(d['ms_per_step'],4), 'raster', round(r['stage_ms_per_step']['raster'],4), 'parity', d['parity_checked'])"
