run() { env WIGNERD_B200_LIB=$PWD/wignerd/jl_b200/libwd_exp.so "$@" timeout 120 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/tmp_b.log 2>&1
  python -c "
import json,sys
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('$*', '%.3f ms' % l['ms_per_step'], flush=True)"; }
WIGNERD_B200_LIB=$PWD/wignerd/jl_b200/libwd_exp.so timeout 300 python tools/quick_check.py > gpurun_out/qc.log 2>&1; echo qc rc=$?; tail -1 gpurun_out/qc.log
run A=1
run WD_K2_NOSTORE=1
for h in 0 103 105 107; do run WD_K2_HANDOFF=$h; done
