# data-parallel bench under different NCCL algorithm / protocol choices (bash tools/nccl_variants.sh N)
N=${1:-4}
port=29520
for v in "NCCL_PROTO=LL,LL128,Simple" "NCCL_PROTO=Simple" "NCCL_ALGO=NVLS" "NCCL_PROTO=LL128"; do
  port=$((port+1))
  env $v NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/nccl_$N.json 2> gpurun_out/nccl_$N.err
  python -c "
import json; d=json.loads(open('gpurun_out/nccl_$N.json').read().strip().splitlines()[-1]); print('$v', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,3), 'M rays/s')" 2>/dev/null || echo "$v failed: $(tail -2 gpurun_out/nccl_$N.err)"
  grep -i -m3 "nvls\|Using network\|comm 0x.* rank 0 .*nranks" gpurun_out/nccl_$N.err | cut -c1-200
  grep -c "NVLS" gpurun_out/nccl_$N.err
done
