# Development: config-5 arm (592 genes, 1 GPU) under different settings; one line each.
O=gpurun_out
BASE="--genes 16 --steps 1 --warmup 1 --no-cpu-baseline --no-dense-arm --no-ss-arm --no-e2e --no-api-arm --config5-genes 592"
i=0
for cfg in "3 16 64" "2 16 64" "4 16 64" "3 32 96"; do
  set -- $cfg
  i=$((i+1))
  SUSHIE_B200_TRACE_HOST=1 python bench.py $BASE --c5-workers $1 --c5-min-items $2 --c5-max-items $3 > $O/c5_sweep_$i.json 2> $O/c5_sweep_$i.err
  python - "$cfg" $O/c5_sweep_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])["config5"]
print("workers/min/max", sys.argv[1], "-> %.2f genes/s  %.1f fits/s  chunks %d" % (d["genes_per_sec"], d["fits_per_sec"], d["chunks"]))
PY
done
grep "chunk of 384" $O/c5_sweep_1.err | grep -v results | head -12
