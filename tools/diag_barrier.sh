# What the per-step band barrier costs: the configs[3] band split at N ranks with the barrier, without it
# (diagnostic: DTR_BENCH_DIAG_NO_BARRIER=1) and with the NCCL gather.  usage: bash tools/diag_barrier.sh <N>
cd /root/repo
N=${1:-2}
run() { # label env gatherflag
env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload fill4k --steps 100 --warmup 5 --no-cpu-baseline --no-others --e2e-steps 1 $3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'N=$N ms/step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms_per_step'].items()}, d.get('parity_checked'))"
}
run barrier DTR_X=0 ""
run nobarrier DTR_BENCH_DIAG_NO_BARRIER=1 ""
run ncclgather DTR_X=0 "--gather nccl"
