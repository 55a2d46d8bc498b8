# per-stage table (ours + reference) and the single-GPU run of the oversize frame.  usage: bash tools/gpu_stages.sh <tag>
TAG=${1:-stages}
mkdir -p gpurun_out
python tools/bench_stages.py --impl ours > gpurun_out/stages_ours_$TAG.jsonl 2> gpurun_out/stages_ours_$TAG.err; tail -3 gpurun_out/stages_ours_$TAG.err
python tools/bench_stages.py --impl reference > gpurun_out/stages_ref_$TAG.jsonl 2> gpurun_out/stages_ref_$TAG.err; tail -3 gpurun_out/stages_ref_$TAG.err
python tools/bench_tiled.py --steps 2 --warmup 1 > gpurun_out/tiled_n1_$TAG.json 2> gpurun_out/tiled_n1_$TAG.err; tail -3 gpurun_out/tiled_n1_$TAG.err; cat gpurun_out/tiled_n1_$TAG.json
python - <<P
import json
def load(f):
    return {(d['config'], d['op']): d for d in map(json.loads, open(f)) if 'op' in d}
a, b = load('gpurun_out/stages_ours_$TAG.jsonl'), load('gpurun_out/stages_ref_$TAG.jsonl')
for k, d in a.items():
    r = b.get(k)
    print(f"{k[0]:13s} {k[1][:58]:58s} {d.get('ms', -1):8.3f} ms {str(d.get('frac_of_measured_hbm')):>7s}  ref {r.get('ms', -1) if r else -1:8.3f} ms", d.get('error', '')[:80])
for k, r in b.items():
    if k not in a: print('ref only', k, r.get('ms'), r.get('error', '')[:100])
P
