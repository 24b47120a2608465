ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_update|k_build' -s 9 -c 4 -f -o gpurun_out/r01_ba_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dt_track_level|k_fast_score|k_match|k_dt_pointcloud' -c 8 -f -o gpurun_out/r01_frontend_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out | tail -4
