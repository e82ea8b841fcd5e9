# The stated multi-GPU configurations on one 8-GPU box (charged 8x: keep it short): bench at N = 8, C4 through the drop-in
# executable with --workers 8/4/2/1, C5 as a queue of mixed files.  Usage (through gpurun --gpus 8): bash tools/gpu_configs8.sh <tag>
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
nvidia-smi -L | head -8; nproc; df -h /dev/shm | tail -1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps ${STEPS:-10} --no-cpu-baseline --no-bd-rate > gpurun_out/${TAG}_scale_n8.json 2> gpurun_out/${TAG}_scale_n8.err
tail -c 300 gpurun_out/${TAG}_scale_n8.err; python -c "
import json
for l in open('gpurun_out/${TAG}_scale_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8', d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms_per_step'])
"
AV1B_CLI_TIMING=1 timeout 600 python tools/c4_run.py --frames ${C4_FRAMES:-2400} --workers 8,4,2,1 --out gpurun_out/${TAG}_c4.json > gpurun_out/${TAG}_c4.log 2>&1; tail -5 gpurun_out/${TAG}_c4.log | cut -c1-700
timeout 400 python tools/queue_bench.py --files 16 --jobs 8 --frames-1080p 96 --frames-4k 32 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err; tail -c 300 gpurun_out/${TAG}_c5.json
