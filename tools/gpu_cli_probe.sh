# CLI tests + a timing probe of the drop-in executable on one GPU.  Usage (through gpurun): bash tools/gpu_cli_probe.sh <tag> [frames]
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1
timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q > gpurun_out/pytest_cli_$TAG.log 2>&1; tail -5 gpurun_out/pytest_cli_$TAG.log
AV1B_CLI_TIMING=1 timeout 400 python tools/c4_run.py --frames ${2:-600} --workers 1 --out gpurun_out/${TAG}_c4_probe.json > gpurun_out/${TAG}_c4_probe.log 2>&1; tail -c 1200 gpurun_out/${TAG}_c4_probe.log
