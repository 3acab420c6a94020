python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --storage i8 --genes 592 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 i8"
python bench.py --storage i8 --genes 1000 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | python scripts/bench_line.py "conv1000 i8 e2e"
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/prof_i8_r1h python bench.py --storage i8 --single-phase --genes 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 4 --min-tol 0 2>&1 | tail -1 | cut -c1-100
