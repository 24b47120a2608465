#!/bin/bash
# Round-end measurement job (run on a B200 box: gpurun --timeout 1500 -- 'bash scripts/profile_job.sh').
# Writes the ncu launch list and full captures plus the bench lines into gpurun_out/; summarise with
#   python scripts/summarize_ncu.py r01 gpurun_out/r01_launches.csv gpurun_out/r01_ba_kernels.ncu-rep \
#          gpurun_out/r01_frontend_kernels.ncu-rep gpurun_out/r01_prep_kernels.ncu-rep
cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_update|k_build|k_regroup' -s 9 -c 4 -f -o gpurun_out/r01_ba_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dt_track_level|k_fast_score|k_fast_select|k_match|k_pose_lm|k_dt_pointcloud' -c 10 -f -o gpurun_out/r01_frontend_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pyrdown_f32|k_pyrdown_u8|k_deriv|k_u8_to_f32' -c 4 -f -o gpurun_out/r01_prep_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
