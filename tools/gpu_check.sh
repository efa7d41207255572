# One-call GPU check used during development: `gpurun -- bash tools/gpu_check.sh`
# (parity suite, then the device-resident bench line of every BASELINE workload).
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('mesh1080', 'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'frac',round(r['frac'],3), r['stage_ms_per_step'])"
for w in mesh4k_tex views1080_tex; do
python bench.py --workload $w --views 16 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', 'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'frac',round(r['frac'],3), r['stage_ms_per_step'])"
done
python bench.py --workload fill4k --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fill4k', 'value',round(d['value'],2),'ms/step',round(d['ms_per_step'],3),'frac',round(r['frac'],3), r.get('stage_ms_per_step'))"
