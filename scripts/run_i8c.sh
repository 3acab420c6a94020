for lib in libsushie_b200.so lib_b2.so; do
SUSHIE_B200_LIB=$PWD/sushie_b200/_lib/$lib python bench.py --storage i8 --genes 592 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-dense-arm --max-iter 50 --min-tol 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel'])"
done
