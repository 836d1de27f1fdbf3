python ba_batch_prof.py > gpurun_out/plain_bab.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_bab.csv python ba_batch_prof.py > gpurun_out/ncu_bab.log 2>&1
echo "batched launch list exit $?"
