#!/bin/bash
cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_update|k_build' -s 9 -c 4 -f -o gpurun_out/r01_ba_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dt_track_level|k_fast_score|k_match|k_dt_pointcloud|k_pose_lm|k_pyrdown_f32|k_deriv' -c 12 -f -o gpurun_out/r01_frontend_kernels python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err
cut -c1-700 gpurun_out/bench_ref_final.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la gpurun_out | tail -5
