#!/bin/bash
# One gpurun session: GPU parity tests, a bench line, the ncu launch list and full captures of the top kernels.
# usage: tools/gpu_session.sh <tag> [tests|bench|ncu ...]
set -u
TAG=${1:-rXX}; shift || true
WHAT=${*:-tests bench ncu}
OUT=gpurun_out; mkdir -p $OUT
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log ;;
bench)
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -c 1500 $OUT/${TAG}_bench.json ;;
ncusplit)
  timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_dp_ -c 8 --csv --log-file $OUT/${TAG}_dp_split.csv \
     python bench.py --steps 1 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_split.log 2>&1; echo "ncu split exit $?"
  for k in k_dp_fill_h k_dp_tail_h; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o $OUT/${TAG}_$k \
       python bench.py --steps 1 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_$k.log 2>&1; echo "ncu $k exit $?"
  done ;;
benchsmall)
  timeout 600 python bench.py --genome-mbp 240 --reads 400000 --batch 200000 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_benchsmall.json 2> $OUT/${TAG}_benchsmall.err; echo "benchsmall exit $?"; tail -c 1200 $OUT/${TAG}_benchsmall.json; tail -5 $OUT/${TAG}_benchsmall.err ;;
benchse)
  timeout 900 python bench.py --workload se100 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_benchse.json 2> $OUT/${TAG}_benchse.err; echo "benchse exit $?"; tail -c 900 $OUT/${TAG}_benchse.json ;;
benchk)
  for k in 0 14 16; do
    timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --seed-table $k > $OUT/${TAG}_benchk$k.json 2> $OUT/${TAG}_benchk$k.err; echo "benchk$k exit $?"
    python -c "import json; d=json.loads(open('$OUT/${TAG}_benchk$k.json').read().strip().splitlines()[-1]); print('k=$k', round(d['value'],2), round(d['e2e']['value'],2), {a:round(b,1) for a,b in d['stage_ms'].items()}, d['work_per_step']['seed_sides'], d['config']['seed_table_build_s'])"
  done ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log ;;
benchref)
  timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "benchref exit $?"; tail -c 1200 $OUT/${TAG}_bench_ref.json; tail -8 $OUT/${TAG}_bench_ref.err ;;
ncufinal)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
  for k in k_seed_search3 k_exact_sweep2; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$k \
       python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_$k.log 2>&1; echo "ncu $k exit $?"
  done ;;
bench2)
  BT2G_DP_PACKED=2 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench2.json 2> $OUT/${TAG}_bench2.err; echo "bench2 exit $?"; tail -c 700 $OUT/${TAG}_bench2.json ;;
bench0)
  BT2G_DP_PACKED=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench0.json 2> $OUT/${TAG}_bench0.err; echo "bench0 exit $?"; tail -c 700 $OUT/${TAG}_bench0.json ;;
ncudp)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_dp_e2e -s 1 -c 1 -f -o $OUT/${TAG}_k_dp_e2e \
     python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_k_dp_e2e.log 2>&1; echo "ncu dp exit $?" ;;
benchfull)
  timeout 1500 python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "bench exit $?"; tail -c 1500 $OUT/${TAG}_bench_full.json ;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
  for k in k_dp_e2e k_seed_search2 k_exact_sweep2; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$k \
       python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > $OUT/${TAG}_ncu_$k.log 2>&1; echo "ncu $k exit $?"
  done ;;
esac
done
