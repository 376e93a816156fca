# round-end validation driver (one gpurun call, one GPU): smoke, the whole GPU test-suite, the default bench, one ncu --set full
# capture of the complex-D kernel -- each under its own timeout
mkdir -p gpurun_out
timeout 150 python __graft_entry__.py --smoke > gpurun_out/smoke_r2.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/smoke_r2.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 gpurun_out/pytest_gpu_r2.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"
bash tools/ncu_run.sh k3_tma cfg3 16000 prof_r2_k3_tma_final --variant 7
