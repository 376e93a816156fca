# kernel-experiment driver: each line = env settings for one short cfg-2 bench run
run() { timeout 300 env "$@" python bench.py --config cfg2 --steps 20 --warmup 3 --no-cpu-baseline --e2e-angles 256 > gpurun_out/b_tmp.log 2>&1; python -c "
import json,sys
l=json.loads(open('gpurun_out/b_tmp.log').read().strip().splitlines()[-1]); print(sys.argv[1:], l['ms_per_step'])" "$@"; }
timeout 200 python tools/quick_check.py 2>&1 | tail -14 | cut -c1-40 | tr '\n' ';'; echo
run WD_X=0
run WD_K2_HANDOFF=101
run WD_K2_HANDOFF=102
run WD_K2_HANDOFF=103
run WD_K2_HANDOFF=104
