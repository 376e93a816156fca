# kernel experiments: several builds of the library (WIGNERD_B200_LIB) x env knobs, cfg-2 timing for paired (variant 4) and unpaired (variant 5)
run() { # lib variant extra-env...
  lib=$1; v=$2; shift 2
  env WIGNERD_B200_LIB=$PWD/wignerd/jl_b200/$lib "$@" timeout 120 python bench.py --config cfg2 --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-extras > gpurun_out/tmp_b.log 2>&1
  python - "$lib" "$v" "$*" <<'PY'
import json,sys
try:
    l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print(sys.argv[1], 'variant', sys.argv[2], sys.argv[3], '%.3f ms' % l['ms_per_step'], flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], sys.argv[3], 'FAILED', open('gpurun_out/tmp_b.log').read()[-300:])
PY
}
for L in $LIBS; do
  run $L 4; run $L 5; run $L 4 WD_K2_NOSTORE=1; run $L 5 WD_K2_NOSTORE=1
done
for L in $LIBS; do run $L 4; run $L 2; done
