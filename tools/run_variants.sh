# scheduling variants of the train step (bench.py value / e2e ms per step)
for v in "PS_FIELD_CHUNKS=3" "PS_FIELD_CHUNKS=2" "PS_FIELD_CHUNKS=4" "PS_FIELD_CHUNKS=6" "PS_OVERLAP_PROP_BWD=0"; do
  env $v python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done
