# guarded experiment driver for the complex-D tensor-store kernel (k3_tma_kernel): parity first, then timing at cfg-3 of
# the shipped library (variant 0 = k3_tma_kernel, variant 2 = k2_cplx_kernel) and of the experiment build (knobs)
mkdir -p gpurun_out
timeout 420 python tools/quick_check.py > gpurun_out/qc.log 2>&1; rc=$?; echo "quick_check rc=$rc"; grep -v "e-1[3-9]\|e-2\| 0.0$" gpurun_out/qc.log | tail -8
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
run() { tag=$1; shift; env "$@" timeout 150 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras $BARGS > gpurun_out/tmp_b.log 2>&1
  python -c "
import json,sys
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('$tag', '%.3f ms' % l['ms_per_step'], l['clocks'].get('sm_mhz'), flush=True)" || tail -3 gpurun_out/tmp_b.log; }
BARGS="--variant 0" run "k3t shipped" A=1
X=$PWD/wignerd/jl_b200/libwd_exp.so
V1=$PWD/wignerd/jl_b200/libwd_v1exp.so
BARGS="--variant 7" run "k3t v1 (first version)" WIGNERD_B200_LIB=$V1
BARGS="--variant 7" run "k3t exp 8 warps" WIGNERD_B200_LIB=$X WD_K3T_NW=8
BARGS="--variant 7" run "k3t exp 12 warps nostore" WIGNERD_B200_LIB=$X WD_K2_NOSTORE=1
for h in ${HANDOFFS:-0 2 3 4 101 102 104}; do BARGS="--variant 7" run "k3t exp 12 warps handoff $h" WIGNERD_B200_LIB=$X WD_K2_HANDOFF=$h; done
if [ -n "$NCU" ]; then bash tools/ncu_run.sh k3_tma cfg3 16000 prof_r2_k3_tma --variant 7; fi
