#!/bin/bash
# Round capture on the GPU box (run under gpurun from the repo root):
#   bash profiles/capture.sh <tag> [tests]
# Writes gpurun_out/<tag>_*: the GPU test log, the bench line, the ncu launch list of the bench command and
# one `ncu --set full` capture per dominant kernel (raw-page CSV extracted on the box).  Numbers printed under
# ncu are never bench values.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
PY=python
if [ "$2" = "tests" ]; then
  timeout 1500 $PY -m pytest tests -m gpu -x -q > $OUT/${TAG}_gputests.log 2>&1
  tail -3 $OUT/${TAG}_gputests.log
fi
timeout 600 $PY bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
# launch list of the same command (short)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  $PY bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
$PY profiles/launchsum.py $OUT/${TAG}_launches_bench.csv | tee $OUT/${TAG}_launch_summary.txt
METRICS='dram__bytes_(read|write)\.sum|dram__cycles_active|gpu__dram_throughput|dram__throughput|sm__pipe_tensor_cycles_active|sm__warps_active|launch__registers_per_thread|gpu__time_duration|sm__inst_executed.sum |smsp__issue_active|sm__throughput|launch__grid_size|launch__block_size|launch__occupancy_limit|l1tex__t_sectors_pipe_lsu_mem_global_op_ld|smsp__inst_executed.sum|launch__waves_per_multiprocessor|sm__inst_issued'
for K in k_denoiser_tc k_score_stream; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${K}\$" -s 3 -c 2 -f -o $OUT/${TAG}_${K} \
    $PY bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --eager > $OUT/${TAG}_ncu_${K}.log 2>&1
  ncu -i $OUT/${TAG}_${K}.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/rawpick.py "$METRICS" > $OUT/${TAG}_${K}_metrics.txt
done
# the split-operand (f16x3) engine and the CTA-pair bf16 engine: one sampler launch each
for E in 3 2; do
  KN=k_denoiser_tc$E
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${KN}\$" -s 2 -c 1 -f -o $OUT/${TAG}_${KN} \
    $PY tests/bench_configs.py sampler 1024 $E > $OUT/${TAG}_ncu_${KN}.log 2>&1
  ncu -i $OUT/${TAG}_${KN}.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/rawpick.py "$METRICS" > $OUT/${TAG}_${KN}_metrics.txt
done
for CELL in "20 8" "200 64"; do
  set -- $CELL
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_dense -s 3 -c 1 -f -o $OUT/${TAG}_dense_T$1_K$2 \
    $PY tests/bench_configs.py dense 262144 $1 $2 > $OUT/${TAG}_ncu_dense_$1_$2.log 2>&1
  ncu -i $OUT/${TAG}_dense_T$1_K$2.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/rawpick.py "$METRICS" > $OUT/${TAG}_dense_T$1_K$2_metrics.txt
done
ls -la $OUT | head -40
