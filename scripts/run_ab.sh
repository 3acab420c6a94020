for i in 1 2; do
for lib in lib_old_k3.so libsushie_b200.so; do
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/$lib python bench.py --genes 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 $lib"
done
done
for lib in lib_old_k3.so libsushie_b200.so; do
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/$lib python bench.py --genes 592 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | python scripts/bench_line.py "conv592 $lib"
done
