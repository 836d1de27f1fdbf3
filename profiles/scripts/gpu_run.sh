set -x
python profiles/scripts/ba_batch_prof.py > gpurun_out/plain_bab.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ba_build_dense_kernel|ba_backsub_kernel' -c 3 -f -o gpurun_out/r02_ba_batch_full python profiles/scripts/ba_batch_prof.py > gpurun_out/ncu_bab_full.log 2>&1
tail -2 gpurun_out/ncu_bab_full.log
