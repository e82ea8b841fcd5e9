# The stated multi-GPU configurations on one 8-GPU box (charged 8x: keep it short): bench at N = 8 (and N = 2), C4 through the
# drop-in executable with --workers 1/2/4/8, C5 as a queue of 32 mixed files.  Usage (through gpurun --gpus 8): bash tools/gpu_configs8.sh <tag>
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
nvidia-smi -L | head -8; nproc; df -h /dev/shm | tail -1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --no-cpu-baseline --no-bd-rate > gpurun_out/${TAG}_scale_n$n.json 2> gpurun_out/${TAG}_scale_n$n.err
  tail -c 400 gpurun_out/${TAG}_scale_n$n.err
done
timeout 900 python tools/c4_run.py --frames ${C4_FRAMES:-1200} --workers 8,4,1 --out gpurun_out/${TAG}_c4.json > gpurun_out/${TAG}_c4.log 2>&1; tail -3 gpurun_out/${TAG}_c4.log | cut -c1-600
timeout 600 python tools/queue_bench.py --files 32 --jobs 8 --frames-1080p 96 --frames-4k 32 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err; tail -c 600 gpurun_out/${TAG}_c5.json
