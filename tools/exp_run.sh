# kernel-experiment driver: each line = env settings for one short bench run
run() { cfgname=$1; shift; timeout 300 env "$@" python bench.py --config $cfgname --steps 20 --warmup 3 --no-cpu-baseline --e2e-angles 256 > gpurun_out/b_tmp.log 2>&1; python -c "
import json,sys
l=json.loads(open('gpurun_out/b_tmp.log').read().strip().splitlines()[-1]); print(sys.argv[1:], l['ms_per_step'])" $cfgname "$@"; }
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_wave or D_batch or ref_wignerD or variants" 2>&1 | tail -3
run cfg3 WD_X=0
run cfg3 WD_K2_HANDOFF=101
run cfg3 WD_K2_HANDOFF=102
run cfg3 WD_K2_HANDOFF=104
run cfg3 WD_K2_HANDOFF=0
