# guarded experiment driver: quick parity check first (short timeout), then the streaming-kernel tests and cfg-4
timeout 90 python tools/quick_check.py > gpurun_out/qc.log 2>&1; rc=$?; echo "quick_check rc=$rc"; tail -5 gpurun_out/qc.log | cut -c1-60
if [ $rc -ne 0 ]; then echo "stopping"; exit 1; fi
timeout 100 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/b_cfg4.log 2>&1; echo "cfg4 rc=$?"; python -c "
import json
l=json.loads(open('gpurun_out/b_cfg4.log').read().strip().splitlines()[-1]); print('cfg4', l['ms_per_step'], l['fp64'])"
timeout 200 python -m pytest tests -m gpu -x -q -k "streaming or large_j or multi_j or very_large or empty_and_single or device_memory or cache_clear" 2>&1 | tail -3
