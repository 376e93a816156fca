# one-shot: parity of the complex-D path, cfg-3 timing, and -- only if it beats LIMIT ms -- the default bench line and an ncu capture
mkdir -p gpurun_out
timeout 100 python tools/quick_check.py > gpurun_out/qc.log 2>&1; rc=$?; echo "quick_check rc=$rc"; grep -v "e-1[3-9]\|e-2\| 0.0$" gpurun_out/qc.log | tail -4
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
timeout 120 python -m pytest tests -q -m gpu -k "tensor_store or D_batch or multi_wave or ref_wignerD or cfg3 or complex or smoke or mp_exact" > gpurun_out/pytest_k3t.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -1 gpurun_out/pytest_k3t.log
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
timeout 60 python bench.py --config cfg3 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/tmp_b.log 2>&1
python -c "
import json,sys
l=json.loads(open('gpurun_out/tmp_b.log').read().strip().splitlines()[-1]); print('cfg3 %.3f ms' % l['ms_per_step']); sys.exit(0 if l['ms_per_step'] < ${LIMIT:-3.20} else 1)" || { echo "not faster: stopping"; exit 0; }
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"
bash tools/ncu_run.sh k3_tma cfg3 16000 prof_r2_k3_tma_final --variant 7
