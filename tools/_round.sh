python tools/microbench.py > gpurun_out/microbench_r1j.json 2> gpurun_out/microbench_r1j.err; tail -c 2500 gpurun_out/microbench_r1j.json
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; cut -c1-1200 gpurun_out/bench_r1k.json
