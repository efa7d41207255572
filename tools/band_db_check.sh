cd /root/repo
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload fill4k --steps 100 --warmup 5 --no-cpu-baseline --no-others --e2e-steps 1 2>gpurun_out/db_n$N.err | python -c "
import sys,json
t=sys.stdin.read().strip()
try:
    d=json.loads(t.splitlines()[-1]); print('N=$N ms/step', round(d['ms_per_step'],4), 'one target', d.get('config'), d.get('parity_checked'))
    print(json.dumps({k:d[k] for k in d if k in ('value','ms_per_step','parity')}))
except Exception as e:
    print('FAILED', e, open('gpurun_out/db_n$N.err').read()[-1500:])
"
echo rc=$?
