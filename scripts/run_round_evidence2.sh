# Round-2 closing evidence on one B200 (gpurun): a GPU test subset over the epilogue / queue paths changed last, the
# default bench exactly as the driver runs it, and an ncu launch list of one config-5 chunk (32 genes x 6 fits through
# the public API) -- taken only after the same command has exited 0 without ncu.
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -k "batch or config5 or sharded or cli or selection or purity or row_view" > $O/r02v_gputest_subset.log 2>&1; echo "pytest rc=$?" >> $O/r02v_gputest_subset.log; tail -3 $O/r02v_gputest_subset.log
python bench.py > $O/r02v_bench_1gpu.json 2> $O/r02v_bench_1gpu.err; echo "bench rc=$?"
python scripts/profile_host.py --genes 32 --top 5 > $O/r02v_chunk_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r02v.csv python scripts/profile_host.py --genes 32 --top 5 > $O/ncu_launches_v.log 2>&1
python scripts/summarize_launches.py $O/launches_r02v.csv $O/r02v_cv_launches > /dev/null; rm -f $O/launches_r02v.csv
python scripts/bench_line.py default < $O/r02v_bench_1gpu.json
