bash tools/gpu_round.sh r02i
cd $GRAFT_REPO_ROOT
timeout 900 python tools/c3_run.py --frames 1200 --out gpurun_out/r02i_c3_1080p10_1200.json > gpurun_out/r02i_c3.log 2>&1; tail -c 900 gpurun_out/r02i_c3.log
timeout 300 python tools/c4_run.py --frames 300 --size 1920x1080 --workers 1 --out gpurun_out/r02i_c4_smoke.json > gpurun_out/r02i_c4_smoke.log 2>&1; tail -c 700 gpurun_out/r02i_c4_smoke.log
timeout 300 python tools/queue_bench.py --files 4 --jobs 2 --frames-1080p 48 --frames-4k 16 > gpurun_out/r02i_c5_smoke.json 2> gpurun_out/r02i_c5_smoke.err; tail -c 500 gpurun_out/r02i_c5_smoke.json
