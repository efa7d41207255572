# Round-2 development check on the GPU box: `gpurun --gpus 2 -- bash tools/gpu_r02.sh <tag>`
TAG=${1:-r02x}
O=gpurun_out
mkdir -p $O
python -X faulthandler -m pytest tests -m gpu -q -rs > $O/${TAG}_pytest.log 2>&1; tail -15 $O/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -5 $O/${TAG}_bench.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/${TAG}_bench_n2.err
fi
TAG=$TAG python - <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG", sys.argv[1] if len(sys.argv)>1 else "r02x"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    r=d["roofline"]
    print(f, "value %.1f ms/step %.3f frac %.3f parity %s e2e %.2f" % (d["value"], d["ms_per_step"], r["frac"], d.get("parity_checked"), (d.get("e2e") or {}).get("value", 0)))
    for o in d.get("other_workloads", []):
        print("   ", o["workload"], "value %.1f ms/step %.3f frac %.3f parity %s" % (o["value"], o["ms_per_step"], o["roofline"]["frac"], o.get("parity_checked")))
PY
if [ -n "$VARIANTS" ]; then bash tools/ab_variants.sh $VARIANTS; fi
