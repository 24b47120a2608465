#!/bin/bash
# Round-end measurement job (run on a B200 box: gpurun --timeout 2400 -- 'bash scripts/profile_job.sh r02').
# Writes the ncu launch list and full captures plus the bench lines into gpurun_out/; summarise here with
#   python scripts/summarize_ncu.py r02 gpurun_out/r02_launches.csv gpurun_out/r02_ba_kernels.ncu-rep \
#          gpurun_out/r02_frontend_kernels.ncu-rep
# --cache-control none: the kernels are profiled with the caches as the preceding kernels left them (the in-loop
# traffic), not flushed before every replay.
R=${1:-r02}
QUICK=${2:-}   # "quick": skip the front-end capture (front-end kernels unchanged since the last full job)
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${R}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/${R}_pytest.txt
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --frames 6 > /dev/null 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_solve|k_update|k_build_wave' -s 9 -c 3 -f -o gpurun_out/${R}_ba_kernels python bench.py --steps 1 --warmup 1 --frames 0 > /dev/null 2>&1
[ "$QUICK" = quick ] || ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_dt_track_level|k_fast_score|k_fast_select|k_match|k_pose_lm' -s 10 -c 8 -f -o gpurun_out/${R}_frontend_kernels python bench.py --steps 1 --warmup 1 --frames 6 > /dev/null 2>&1
cat gpurun_out/${R}_pytest.txt; tail -c 600 gpurun_out/${R}_bench.json; ls -la gpurun_out/${R}_*
