run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 12 --warmup 4 --no-extras --no-cpu-baseline 2>gpurun_out/nvar.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['value']))"; }
echo -n "peer-dma: "; PS_EXCHANGE=peer run 29701
echo -n "nccl: "; PS_EXCHANGE=nccl run 29702
