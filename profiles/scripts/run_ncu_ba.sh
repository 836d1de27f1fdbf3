set -x
python profiles/scripts/ba_batch_prof.py > gpurun_out/plain_bab_s2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ba_build_kernel|ba_backsub_kernel|ba_solve_small' -c 4 -f -o gpurun_out/ba_batch_full python profiles/scripts/ba_batch_prof.py > gpurun_out/ncu_bab_full.log 2>&1
python profiles/scripts/ba_large_prof.py > gpurun_out/plain_bal_s2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ba_schur_pairs|ba_cam_rows|ba_backsub_kernel|ba_build_kernel' -c 6 -f -o gpurun_out/ba_large_full python profiles/scripts/ba_large_prof.py > gpurun_out/ncu_bal_full.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_s2.json 2>gpurun_out/bench_ref_s2.err
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_s2b_n1.json 2>gpurun_out/bench_s2b_n1.err
ls -la gpurun_out/*.ncu-rep
python -c "
import json
for f in ('gpurun_out/bench_ref_s2.json','gpurun_out/bench_s2b_n1.json'):
    d=json.load(open(f)); print(f, d['value']/1e9, d.get('cpu_baseline'))
"
