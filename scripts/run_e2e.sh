python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --genes 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-dense-arm 2>&1 | tail -1 > gpurun_out/e2e_batched.json
python -c "
import json; d=json.load(open('gpurun_out/e2e_batched.json')); print('value', d['value'], 'e2e', d['e2e']); print(d['roofline'])"
