run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --nx $2 --steps 300 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 nx$2', d['ms_per_step'], d['rank_phase_ms']['rows'][0])"; }
for nx in 13 20; do run base $nx; TM_P2P_SPLIT=1 run split $nx; run base2 $nx; TM_P2P_SPLIT=1 run split2 $nx; done
