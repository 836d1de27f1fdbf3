python sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/sanitize_memcheck.log
python proj_prof.py > gpurun_out/plain_proj.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_proj.csv python proj_prof.py > gpurun_out/ncu_proj.log 2>&1
echo "proj launch list exit $?"
