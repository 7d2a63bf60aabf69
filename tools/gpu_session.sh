#!/bin/bash
# One gpurun session (round 2: the exact engine is the measured path).
# usage: tools/gpu_session.sh <tag> [tests|smoke|bench|benchse|benchloc|benchl|benchref|launches|ncustep|ncudp|dpx|streamcxx|builder ...]
set -u
TAG=${1:-rXX}; shift || true
WHAT=${*:-tests bench}
OUT=gpurun_out; mkdir -p $OUT
SMALL="--genome-mbp 240 --seed-table 14 --dense-sa 0 --reads 200000 --batch 200000 --engines 1 --steps 1 --warmup 0 --no-cpu-baseline --no-text-e2e"
for w in $WHAT; do
case $w in
tests)    timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log ;;
smoke)    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log ;;
bench)    timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -c 1500 $OUT/${TAG}_bench.json ;;
benchse)  timeout 900 python bench.py --workload se100 > $OUT/${TAG}_bench_se100.json 2> $OUT/${TAG}_bench_se100.err; echo "se100 exit $?" ;;
benchloc) timeout 900 python bench.py --workload loc300 --reads 200000 --batch 100000 --steps 1 --warmup 1 --cpu-sample 50000 --no-text-e2e > $OUT/${TAG}_bench_loc300.json 2> $OUT/${TAG}_bench_loc300.err; echo "loc300 exit $?" ;;
benchl)   timeout 900 python bench.py --workload pe150l --steps 3 --warmup 1 --cpu-sample 200000 > $OUT/${TAG}_bench_pe150l.json 2> $OUT/${TAG}_bench_pe150l.err; echo "pe150l exit $?" ;;
benchref) timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "reference arm exit $?" ;;
launches) # per-launch time + DRAM bytes of one 200 k-pair step on the 3 Gbp index -> tools/ncu_summary.py
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --reads 200000 --batch 200000 --engines 1 --steps 1 --warmup 0 --no-cpu-baseline --no-text-e2e > $OUT/${TAG}_launch_bench.json 2> $OUT/${TAG}_launch_bench.err; echo "ncu launches exit $?" ;;
ncustep)  # (a 51 GB seed table cannot be saved / restored for kernel replay: small index, 14-mer table)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_xe_step --launch-skip 4 -c 2 -f -o $OUT/${TAG}_k_xe_step python bench.py $SMALL > $OUT/${TAG}_ncu_step.log 2>&1; echo "ncu step exit $?" ;;
ncudp)    timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_dp_ --launch-skip 8 -c 2 -f -o $OUT/${TAG}_k_dp python bench.py $SMALL > $OUT/${TAG}_ncu_dp.log 2>&1; echo "ncu dp exit $?" ;;
dpx)      ./tools/dpx_bench > $OUT/${TAG}_dpx.json; cat $OUT/${TAG}_dpx.json ;;
streamcxx) timeout 300 python tools/check_stream_cxx_gpu.py > $OUT/${TAG}_stream_cxx.log 2>&1; echo "stream cxx exit $?"; tail -3 $OUT/${TAG}_stream_cxx.log ;;
builder)  timeout 600 python tools/check_builder_identity.py > $OUT/${TAG}_builder.log 2>&1; echo "builder exit $?"; tail -1 $OUT/${TAG}_builder.log ;;
esac
done
