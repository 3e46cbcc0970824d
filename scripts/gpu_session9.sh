set -x
python scripts/anderson_probe.py > gpurun_out/anderson.log 2>&1; cat gpurun_out/anderson.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_s9.log 2>&1; tail -12 gpurun_out/pytest_s9.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-api-leg > gpurun_out/bench_s9.log 2>&1; grep '^{"metric"' gpurun_out/bench_s9.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.2f e2e %.2f steady %.2f launches %d/%d' % (d['value'], d['e2e']['value'], d['steady_state']['value'], d['gpu_launches'], d['steady_state']['gpu_launches']))
print('cold', {k: round(v,2) for k,v in d['stages_ms'].items()}); print('steady', {k: round(v,2) for k,v in d['steady_state']['stages_ms'].items()})
print('gate', d['parity_gate'])
"
