# guarded experiment driver for the complex-D tensor-store kernel (k3_tma_kernel): parity first (quick check + the tests that
# touch complex D), then timing at cfg-3 of the shipped library and of the experiment build (knobs WD_K2_NOSTORE, WD_K2_HANDOFF)
mkdir -p gpurun_out
timeout 150 python tools/quick_check.py > gpurun_out/qc.log 2>&1; rc=$?; echo "quick_check rc=$rc"; grep -v "e-1[3-9]\|e-2\| 0.0$" gpurun_out/qc.log | tail -6
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
timeout 200 python -m pytest tests -q -m gpu -k "tensor_store or D_batch or multi_wave or ref_wignerD or cfg3 or complex or smoke" > gpurun_out/pytest_k3t.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_k3t.log
run() { tag=$1; shift; env "$@" timeout 60 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras $BARGS > gpurun_out/tmp_b.log 2>&1
  python -c "
import json,sys
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('$tag', '%.3f ms' % l['ms_per_step'], l['clocks'].get('sm_mhz'), flush=True)" || tail -3 gpurun_out/tmp_b.log; }
BARGS="--variant 0" run "k3t shipped" A=1
X=$PWD/wignerd/jl_b200/libwd_exp.so
BARGS="--variant 7" run "k3t exp nostore" WIGNERD_B200_LIB=$X WD_K2_NOSTORE=1
