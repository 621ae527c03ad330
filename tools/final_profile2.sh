set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest_gpu.log 2>&1; tail -2 gpurun_out/f_pytest_gpu.log
for RX in k_path_walk k_run_finish; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 2 -c 3 -o gpurun_out/ffull2_${RX} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/f_ncu_full2_${RX}.log 2>&1
done
python bench.py --workload c5 > gpurun_out/f_bench_c5.json 2> gpurun_out/f_c5.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(k_|Device)" --csv --log-file gpurun_out/f_c5_launches.csv \
  python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/f_c5_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/f_c5_launches.csv > gpurun_out/f_c5_launches.txt; head -16 gpurun_out/f_c5_launches.txt
