# parity check of the small channel case on 4 processes (block_dims 2,1,2), several times (races show as run-to-run variation)
run() { name=$1; shift; echo "== $name"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 4 --grid 64 32 64 --steps 2 --warmup 1 --no-cpu --no-alt --e2e-steps 1 $EXTRA 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); c=d.get('check'); print(c['rel_l2'], c['per_field'], c['ok'])
"; }
for i in 1 2 3; do EXTRA="" run base$i A=1; done
EXTRA="--solver mg" run mg A=1
