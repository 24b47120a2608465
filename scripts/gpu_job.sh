python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -2 gpurun_out/bench_b.err; cat gpurun_out/bench_b.json
