for lib in libsushie_b200.so lib_b1_k1.so; do
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/$lib python bench.py --storage i8 --genes 592 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 i8 $lib"
done
