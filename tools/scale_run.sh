#!/bin/bash
# Developer tool: 1/2/4/8-GPU strong-scaling table of bench.py on one box (run under gpurun --gpus 8).
#   tools/scale_run.sh <tag> [extra bench args]      -> gpurun_out/scale_<tag>.jsonl
tag=$1; shift
out=gpurun_out/scale_$tag.jsonl
: > $out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep "$@" >> $out 2>> gpurun_out/scale_$tag.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@" >> $out 2>> gpurun_out/scale_$tag.err
  fi
done
python - "$out" <<'PY'
import json,sys
rows=[json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith('{')]
t1=None
for d in rows:
    if d['n_gpus']==1: t1=d['ms_per_step']
    eff = t1/(d['n_gpus']*d['ms_per_step']) if t1 else float('nan')
    print('N=%d step %.3f ms eff %.3f | e2e %.3f ms | parity %s | fwd exchange %s' % (d['n_gpus'], d['ms_per_step'], eff, d['e2e']['ms_per_step'], d.get('parity',{}).get('ok'), d.get('breakdown',{}).get('forward_exchange')))
PY
