mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40) > gpurun_out/r2_t17.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/r2_t17.log | head -30
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
