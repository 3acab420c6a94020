# Round evidence: default bench (as the driver runs it), reference arm, ncu launch list, ncu --set full
# captures of the three kernels (hard-call, dense fp64, summary-level fp32), config-4 bench.
set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python scripts/bench_ss.py --iters 20 --steps 3 --warmup 3 > gpurun_out/bench_ss.json 2> gpurun_out/bench_ss.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --genes 296 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1c_ind_i8 python bench.py --storage i8 --single-phase --genes 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-dense-arm --max-iter 4 --min-tol 0 > gpurun_out/ncu_i8.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1c_ind_f64 python bench.py --storage f64 --single-phase --genes 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 4 --min-tol 0 > gpurun_out/ncu_f64.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -c 1 -f -o gpurun_out/prof_r1c_ss_f32 python scripts/bench_ss.py --iters 5 --steps 1 --warmup 0 > gpurun_out/ncu_ss.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r1c_ind_i8.ncu-rep gpurun_out/r01_ncu_full_ind_i8.json --loci 148 --iters 4 --bytes-per-ser 3000000 --tag "config 3 shape, hard-call int8 storage, ibss_kernel<3,int8_t,false,1>" > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r1c_ind_f64.ncu-rep gpurun_out/r01_ncu_full_ind_f64.json --loci 296 --iters 4 --bytes-per-ser 24000000 --tag "config 3 shape, dense fp64 storage, ibss_kernel<3,double,false>" > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r1c_ss_f32.ncu-rep gpurun_out/r01_ncu_full_ss_f32.json --loci 1 --iters 5 --first-traversal 0 --bytes-per-ser 1200000000 --tag "config 4, summary level p=10000 fp32 LD, 148-block co-operative team, ibss_kernel<3,float,true>" > /dev/null
rm -f gpurun_out/*.ncu-rep
python scripts/summarize_launches.py gpurun_out/launches_r1c.csv gpurun_out/r01c_launches > /dev/null; rm -f gpurun_out/launches_r1c.csv
python scripts/bench_line.py default < gpurun_out/bench_default.json
cut -c1-330 gpurun_out/bench_ss.json
