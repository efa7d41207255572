# `gpurun --gpus 8 -- bash tools/scale_check.sh r02g` : smoke + the driver's scaling commands at N = 8 (and 4)
TAG=${1:-r02x}; O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke 2>&1 | tail -2
NG=$(nvidia-smi -L | wc -l)
for N in 8 4; do
if [ "$NG" -ge "$N" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" $O/${TAG}_bench_n$N.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/${TAG}_bench_ref_n$N.json 2> $O/${TAG}_bench_ref_n$N.err; echo "ref n$N rc=$?"
fi
done
TAG=$TAG python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_n*.json" % os.environ["TAG"])):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"unparsable",e); continue
    r=d["roofline"]; e=d.get("e2e") or {}
    print(f, "N",d["n_gpus"],"value %.1f ms/step %.3f frac %.3f parity %s e2e %.2f (ceiling frac %s)" % (d["value"], d["ms_per_step"], r["frac"], d.get("parity_checked"), e.get("value",0), e.get("frac_of_d2h_ceiling")))
    for o in d.get("other_workloads", []):
        print("   ", o["workload"], "value %.1f ms/step %.3f frac %.3f parity %s %s" % (o["value"], o["ms_per_step"], o["roofline"]["frac"], o.get("parity_checked"), o["roofline"]["stage_ms_per_step"] if o["workload"]=="fill4k" else ""))
PY
