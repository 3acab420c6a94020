# Round-2 evidence on one B200 (gpurun): GPU test-suite, default bench (as the driver runs it), reference arm,
# ncu launch list of a short bench, ncu --set full capture of the bit-plane IBSS kernel.  Every ncu command runs
# only after the same command line has exited 0 without ncu.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02_gputest.log; tail -3 $O/r02_gputest.log
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
SHORT="--genes 296 --steps 1 --warmup 1 --no-cpu-baseline --no-ss-arm --no-api-arm --config5-genes 0 --dense-steps 3"
python bench.py $SHORT > $O/r02_short.json 2> $O/r02_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_r02.csv python bench.py $SHORT > $O/ncu_launches.log 2>&1
python scripts/summarize_launches.py $O/launches_r02.csv $O/r02_launches > /dev/null; rm -f $O/launches_r02.csv
B2="--storage b2 --single-phase --genes 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-dense-arm --no-ss-arm --no-api-arm --config5-genes 0 --max-iter 4 --min-tol 0"
python bench.py $B2 > $O/r02_b2_plain.json 2> $O/r02_b2_plain.err &&
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o $O/prof_r02_ind_b2 python bench.py $B2 > $O/ncu_b2.log 2>&1
python scripts/ncu_summary.py $O/prof_r02_ind_b2.ncu-rep $O/r02_ncu_full_ind_b2.json --loci 148 --iters 4 --bytes-per-ser 1500000 --tag "config 3 shape, bit-plane hard-call storage, ibss_kernel<3,uint8_t,false,0>" > /dev/null
python scripts/bench_line.py default < $O/r02_bench_1gpu.json
