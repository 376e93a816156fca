# guarded experiment driver of k3_tmap_kernel (12 warps, variant 8): parity under a short timeout, then cfg-3 timing
mkdir -p gpurun_out
timeout 70 python tools/k3p_check.py > gpurun_out/k3p_check.log 2>&1; rc=$?; echo "k3p_check rc=$rc"; tail -4 gpurun_out/k3p_check.log
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
X=$PWD/wignerd/jl_b200/libwd_exp.so
for h in ${HANDOFFS:-103 105 0}; do
  WIGNERD_B200_LIB=$X WD_K2_HANDOFF=$h timeout 45 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --variant 8 > gpurun_out/tmp_b.log 2>&1
  python -c "
import json
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('k3p handoff $h', '%.3f ms' % l['ms_per_step'])" || echo "handoff $h failed"
done
