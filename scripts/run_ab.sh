for lib in lib_old_k3.so lib_direct.so lib_evict.so; do
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/$lib python bench.py --storage f64 --genes 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 f64 $lib"
done
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/lib_evict.so python bench.py --storage i8 --genes 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 i8 lib_evict"
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/lib_evict.so python scripts/bench_ss.py --iters 10 --steps 3 --warmup 2 2>&1 | tail -1 | cut -c1-330
