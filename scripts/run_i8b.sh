for st in i8; do
python bench.py --storage $st --genes 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py "fixed50 $st"
done
