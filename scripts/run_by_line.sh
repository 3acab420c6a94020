# Per-source-line stall tables of an individual-level kernel (ncu --set full + SASS line info).
#   bash scripts/run_by_line.sh <storage: b2|i8|f64> <mangled kernel substring> <out prefix>
set -x
ST=${1:-b2}; KN=${2:-ibss_kernelILi3EhLb0ELi0E}; OUT=${3:-gpurun_out/by_line_$ST}
ARGS="--storage $ST --single-phase --genes 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-dense-arm --no-ss-arm --no-api-arm --max-iter 4 --min-tol 0"
python bench.py $ARGS > $OUT.plain.json 2> $OUT.plain.err &&
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o $OUT python bench.py $ARGS > /dev/null 2>&1
python scripts/ncu_by_line.py $OUT.ncu-rep sushie_b200/_lib/libsushie_b200.so $KN 48 > $OUT.txt 2>&1
python scripts/ncu_summary.py $OUT.ncu-rep $OUT.json --loci 148 --iters 4 --bytes-per-ser 1500000 --tag "config 3 shape, $ST storage" > /dev/null 2>&1
ls -la $OUT.ncu-rep
head -60 $OUT.txt | cut -c1-220
