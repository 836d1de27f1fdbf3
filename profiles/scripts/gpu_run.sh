set -x
timeout 600 python bench.py --workload sweep --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_sweep.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_sweep_launches.csv python bench.py --workload sweep --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_sweep.log 2>&1; echo rc=$?
timeout 600 python bench.py --workload ba_large --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bal.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ba_large_launches.csv python bench.py --workload ba_large --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bal.log 2>&1; echo rc=$?
