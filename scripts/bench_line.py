"""Print a compact one-line summary of a bench.py JSON line read from stdin."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
txt = sys.stdin.read().strip().splitlines()
line = next((l for l in reversed(txt) if l.startswith("{")), None)
if line is None:
    print(tag, "NO JSON:", txt[-3:]); sys.exit(0)
d = json.loads(line)
k = d.get("kernel", {})
print(f"{tag}: {d['value']:.2f} {d['unit']} ms/step={d['ms_per_step']:.1f} frac={d['roofline']['frac']:.4f} team={k.get('team_size')} launches={d.get('gpu_launches')} iters/locus={d.get('mean_iters_per_locus'):.1f} clocks={d['clocks']['sm_mhz']} e2e={d.get('e2e',{}).get('value')}")
