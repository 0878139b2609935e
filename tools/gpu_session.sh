#!/bin/bash
# One gpurun call: GPU parity tests, default bench, ncu launch list, ncu --set full captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> [tests] [bench] [launches] [full:<regex>:<skip>:<count>] ...'
# Outputs land in gpurun_out/<tag>_*.  Numbers printed under ncu are never bench values.
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)
      timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench)
      timeout 600 python bench.py > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench.log ;;
    benchq)
      timeout 600 python bench.py --no-cpu --no-scan --quick > gpurun_out/${tag}_benchq.log 2>gpurun_out/${tag}_benchq.err; echo "benchq rc=$?"; tail -c 2500 gpurun_out/${tag}_benchq.log ;;
    refarm)
      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_refarm.log 2>&1; echo "refarm rc=$?" ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --quick --no-scan \
        > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launches rc=$?" ;;
    traffic)
      # DRAM bytes + time of every GEMM launch of ONE step (warm-up = 3 steps x 144 launches are skipped)
      timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tcgen05 -s 432 -c 144 --csv --log-file gpurun_out/${tag}_gemm_traffic.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu --no-scan --no-graph --quick > gpurun_out/${tag}_ncu_traffic.log 2>&1; echo "traffic rc=$?" ;;
    full:*)
      IFS=: read -r _ rx skip cnt <<< "$what"
      name=$(echo "${rx}_${skip}" | tr -c 'A-Za-z0-9_' '_' | cut -c1-28)
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s ${skip:-0} -c ${cnt:-2} \
        -f -o gpurun_out/${tag}_full_${name} python bench.py --steps 1 --warmup 3 --no-cpu --no-graph --quick --no-scan \
        > gpurun_out/${tag}_ncu_full_${name}.log 2>&1; echo "full $rx rc=$?" ;;
    *) bash -c "$what" ;;
  esac
done
