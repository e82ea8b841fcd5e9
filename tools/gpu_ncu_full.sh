# ncu --set full capture of the hot kernels out of a short bench run (one GPU); the report is read on the box and only the
# CSV pages come home (gpurun_out is capped at 64 MiB).  Usage (through gpurun): bash tools/gpu_ncu_full.sh <tag> [regex] [count]
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
RE=${2:-'inter_encode_kernel|mctf_kernel|hme_refine_kernel|hme_sbrd_kernel|cdef_kernel|hme_l2_kernel|deblock_kernel'}
CNT=${3:-22}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 44 -c $CNT -o /tmp/${TAG}_full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-bd-rate > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
ls -la /tmp/${TAG}_full.ncu-rep
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv > gpurun_out/${TAG}_full_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_full_source.csv
SZ=$(stat -c %s /tmp/${TAG}_full.ncu-rep)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/${TAG}_full.ncu-rep gpurun_out/; fi
ls -la gpurun_out/
