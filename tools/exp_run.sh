# guarded experiment driver: quick parity check first (short timeout), then timing of the kernel variants at cfg-2
timeout 300 python tools/quick_check.py > gpurun_out/qc.log 2>&1; rc=$?; echo "quick_check rc=$rc"; grep -v "e-1[3-9]\|e-2\| 0.0$" gpurun_out/qc.log | tail -5
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
for v in ${VARIANTS:-0 4 5 2}; do
  timeout 120 python bench.py --config cfg2 --steps 20 --warmup 3 --variant $v --no-cpu-baseline --e2e-angles 2048 $BENCH_EXTRA > gpurun_out/b_v$v.log 2>&1; echo "variant $v rc=$?"
  python -c "
import json
l=json.loads(open('gpurun_out/b_v$v.log').read().strip().splitlines()[-1]); print('variant $v', l['ms_per_step'], l['gpu_launches'], l['clocks'])"
done
