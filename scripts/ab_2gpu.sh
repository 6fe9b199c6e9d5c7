run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --nx 13 --steps 300 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['rank_phase_ms']['rows'][0])"; }
run base
TM_GEMM_NO_MULTI=1 run nomulti
TM_GEMM_NO_MULTI_FWD=1 run nomulti_fwd
TM_GEMM_NO_MULTI_BWD=1 run nomulti_bwd
run base2
one() { timeout 100 python bench.py --config $2 --steps 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', d['ms_per_step'])"; }
for c in c3 c5; do one base $c; TM_GEMM_NO_MULTI_FWD=1 one nofwd $c; TM_GEMM_NO_MULTI_BWD=1 one nobwd $c; TM_GEMM_NO_MULTI=1 one nomulti $c; done
