mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60) > gpurun_out/r2_t13.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/r2_t13.log | head -40
