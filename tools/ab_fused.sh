# A/B of the one-kernel form of the opaque-only raster stage (the default: raster_opaque_kernel<true>; DTR_B200_FUSED=0 selects the pair) against
# visibility + resolve, same box: bench.py's default line (512 textured 1080p views) with its other_workloads,
# every number parity checked by bench.py itself.
# usage: gpurun -- bash tools/ab_fused.sh <tag> [variant...]     (variants/libdtr_<variant>.so, run with DTR_B200_FUSED=1)
cd /root/repo
TAG=${1:-fused}; shift
run() { # label fused lib
DTR_B200_LIB=$3 DTR_B200_FUSED=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
echo "== $1 (DTR_B200_FUSED=$2) rc=$?"
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('views1080_tex x512', 'ms/step', round(d['ms_per_step'],4), 'raster', round(r['stage_ms_per_step']['raster'],4), 'iso', round(r['stage_ms_isolated']['raster'],4), 'frac', round(r['frac'],4), 'parity', d['parity_checked'], r.get('raster_kernels_ms_per_step'))
    for o in d['other_workloads']:
        ro=o['roofline']; print(o['workload'], 'ms/step', round(o['ms_per_step'],4), 'raster', round(ro['stage_ms_per_step']['raster'],4), 'frac', round(ro['frac'],4), 'parity', o['parity_checked'])
except Exception as e:
    print('FAILED', e, open('gpurun_out/${TAG}_$1.err').read()[-1500:])
P
}
if [ "${BASE:-1}" != "0" ]; then for f in ${ORDER:-1 0}; do run base_f$f $f /root/repo/dtrenderer_b200/libdtr_b200.so; done; fi
for v in "$@"; do run $v 1 /root/repo/variants/libdtr_$v.so; done
