# experiment: three warps per SM sub-partition with TWO main-loop tokens per ring (k3_tma_kernel, experiment build)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_store or D_batch or multi_wave" > gpurun_out/pytest_k3t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_k3t.log
run() { tag=$1; shift; env "$@" timeout 150 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras $BARGS > gpurun_out/tmp_b.log 2>&1
  python -c "
import json,sys
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('$tag', '%.3f ms' % l['ms_per_step'], l['clocks'].get('sm_mhz'), flush=True)" || tail -3 gpurun_out/tmp_b.log; }
X=$PWD/wignerd/jl_b200/libwd_exp.so
BARGS="--variant 0" run "k3t shipped (8 warps)" A=1
for h in 103 101 3; do
BARGS="--variant 7" run "12 warps 2 tokens handoff $h" WIGNERD_B200_LIB=$X WD_K3T_NW=12 WD_K3T_TOKENS=2 WD_K2_HANDOFF=$h
BARGS="--variant 7" run "12 warps 2 tokens handoff $h nostore" WIGNERD_B200_LIB=$X WD_K3T_NW=12 WD_K3T_TOKENS=2 WD_K2_HANDOFF=$h WD_K2_NOSTORE=1
done
BARGS="--variant 7" run "8 warps nostore" WIGNERD_B200_LIB=$X WD_K2_NOSTORE=1
