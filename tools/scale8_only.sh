# `gpurun --gpus 8 -- bash tools/scale8_only.sh r02n`: the driver's bench command at N = 8 only (no reference arm)
TAG=${1:-r02x}; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; echo "bench n8 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" $O/${TAG}_bench_n8.err | tail -5
TAG=$TAG python - <<'PY'
import json,os
f="gpurun_out/%s_bench_n8.json" % os.environ["TAG"]
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]; e=d.get("e2e") or {}
print("N",d["n_gpus"],"value %.1f ms/step %.3f frac %.3f parity %s e2e %.2f" % (d["value"], d["ms_per_step"], r["frac"], d.get("parity_checked"), e.get("value",0)))
for o in d.get("other_workloads", []):
    print("   ", o["workload"], "value %.1f ms/step %.4f frac %.3f parity %s one-target %s" % (o["value"], o["ms_per_step"], o["roofline"]["frac"], o.get("parity_checked"), o.get("ms_per_step_one_target")))
PY
