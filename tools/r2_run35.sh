O=gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for r in 1 0 1 0; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --attention-ring $r > $O/r2_ring$r.json 2> $O/r2_ring$r.err; python -c "
import json; d=json.load(open('$O/r2_ring$r.json')); print('ring=$r', round(d['value']), d['clocks']['sm_mhz'])"
done
