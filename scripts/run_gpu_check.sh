set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python scripts/bench_ss.py --iters 10 --steps 3 --warmup 2 2>&1 | tail -3
python bench.py --genes 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 50 --min-tol 0 2>&1 | python scripts/bench_line.py fixed50
