# round-end validation driver (one gpurun call, one GPU): smoke, the whole GPU test-suite, the ragged stress run, the default
# bench, then the ncu launch list of the bench command and one --set full capture each of the two real-d kernels
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke_r2.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/smoke_r2.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 gpurun_out/pytest_gpu_r2.log
if [ $rc -ne 0 ]; then echo "pytest failed: stopping"; exit 1; fi
WD_SANITIZE_BIG=1 timeout 600 python tools/sanitize_case.py > gpurun_out/r2_ragged_stress.log 2>&1; echo "ragged stress rc=$?"; tail -1 gpurun_out/r2_ragged_stress.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_ref.log 2>&1; echo "bench ref rc=$?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_plain_list.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
bash tools/ncu_run.sh k2_col cfg2 40000 prof_r2_k2_col_final --no-extras
bash tools/ncu_run.sh k2_real cfg2 16000 prof_r2_k2_real_final --no-extras --grid random
