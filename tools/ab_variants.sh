# A/B of kernel variants on the GPU box: build variants/libdtr_<name>.so (same sources, other -D flags), then
# `gpurun -- bash tools/ab_variants.sh <name>...`; the library under test is selected with DTR_B200_LIB.
cd /root/repo
run() { # name lib
  for w in mesh1080 fill4k; do
  DTR_B200_LIB=$2 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $w', 'value',round(d['value'],2),'ms/step',round(d['ms_per_step'],4),'raster',round(r['stage_ms_per_step']['raster'],4), 'iso', round(r.get('stage_ms_isolated',{}).get('raster',0),4))"
  done
}
run base /root/repo/dtrenderer_b200/libdtr_b200.so
for v in "$@"; do run $v /root/repo/variants/libdtr_$v.so; done
run base2 /root/repo/dtrenderer_b200/libdtr_b200.so
