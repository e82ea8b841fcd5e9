# GPU round + a timing probe of the drop-in executable (AV1B_CLI_TIMING=1).  Usage (through gpurun): bash tools/gpu_round_cli.sh <tag>
TAG=${1:-r02x}
bash tools/gpu_round.sh $TAG
cd $GRAFT_REPO_ROOT
AV1B_CLI_TIMING=1 timeout 400 python tools/c4_run.py --frames 600 --workers 1 --out gpurun_out/${TAG}_c4_probe.json > gpurun_out/${TAG}_c4_probe.log 2>&1; tail -c 1500 gpurun_out/${TAG}_c4_probe.log
