#!/bin/bash
# Dev tool: device / e2e / e2e_json time of bench.py's workload against the pipeline knobs.
for k in 2 4 8; do for n in 1 2 4; do
export KA_PIPELINE_STAGES=$k KA_CHAIN_SUBBLOCKS=$n
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('K=%s nsub=%s value %.3g ms %.3f | e2e %.3g ms %.3f | json ms %.3f' % (os.environ['KA_PIPELINE_STAGES'], os.environ['KA_CHAIN_SUBBLOCKS'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_json']['ms_per_step']))"
done; done
