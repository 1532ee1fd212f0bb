python tools/timeline.py > gpurun_out/timeline_l.txt 2> gpurun_out/timeline_l.err
for v in "PS_OVERLAP_PROP_BWD=0" "PS_FIELD_CHUNKS=2" "PS_FIELD_CHUNKS=4" "PS_FIELD_CHUNKS=1" "PS_FIELD_CHUNKS=1 PS_OVERLAP_PROP_BWD=0" "PS_FIELD_CHUNKS=6"; do
  env $v python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done
