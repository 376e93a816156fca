# round-end validation driver (one gpurun call): smoke first (abort on failure), the whole GPU test-suite, cfg-4, ncu of the streaming kernel
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/smoke_final.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 gpurun_out/pytest_gpu_final.log
if [ $rc -ne 0 ]; then echo "pytest failed: stopping"; exit 1; fi
timeout 100 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_final.log 2>&1; echo "cfg4 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k2_stream -s 300 -c 1 -o gpurun_out/prof_r1_k2_stream_final -f python bench.py --config cfg4 --nb 512 --steps 1 --warmup 3 > gpurun_out/ncu_stream_f.log 2>&1; echo "ncu stream rc=$?"
