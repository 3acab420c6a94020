set -x
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -c 1 -f -o gpurun_out/prof_ss_r1c python scripts/bench_ss.py --iters 5 --steps 1 --warmup 0 2>&1 | tail -5
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/prof_ind_r1c python bench.py --genes 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 4 --min-tol 0 2>&1 | tail -3
