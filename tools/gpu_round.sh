# One GPU round trip: smoke, GPU parity tests, the default bench line, then (optionally) the ncu launch list of the same command.
# Usage (through gpurun): bash tools/gpu_round.sh <tag> [ncu]
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
if [ "$2" = "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-bd-rate > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
fi
tail -3 gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -c 300 gpurun_out/bench_$TAG.json
