set -x
nvidia-smi -L | head -8; nproc
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  tail -c 600 gpurun_out/scale_n$n.err; python -c "
import json,sys
for l in open('gpurun_out/scale_n$n.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print($n, d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['host_threads'], d['breakdown_ms_per_step'])
"
done
