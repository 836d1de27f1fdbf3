# large-BA regression + timing pass
python -m pytest tests/test_ba_gpu.py tests/test_edge_cases_gpu.py tests/test_ref_golden_gpu.py -x -q -m gpu 2>&1 | tail -2
python bench.py --workload ba_large --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_ba_large_q.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_ba_large_q.json'))
print('ba_large: %.1f M obs*iter/s; ms/step %.3f' % (d['value'] / 1e6, d['ms_per_step']), {k: d[k] for k in d if 'iter' in k or 'rmse' in k})
PY
python profiles/scripts/ba_large_prof.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bal_q.csv python profiles/scripts/ba_large_prof.py > /dev/null 2>&1
python profiles/scripts/launch_summary.py gpurun_out/launches_bal_q.csv
