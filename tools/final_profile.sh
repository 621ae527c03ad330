set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/f_bench1024_n1.json 2> gpurun_out/f_bench1024_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench1024_reference_arm.json 2> gpurun_out/f_ref.err
python bench.py --workload c2 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_c2.err
python bench.py --workload c4 > gpurun_out/f_bench_c4.json 2> gpurun_out/f_c4.err
python bench.py --workload c5 > gpurun_out/f_bench_c5.json 2> gpurun_out/f_c5.err
python bench.py --order 5 --no-e2e --no-cpu > gpurun_out/f_bench_1024_order5.json 2> gpurun_out/f_o5.err
KF='regex:^(k_|Device)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/f_launches_1024.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/f_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/f_launches_1024.csv > gpurun_out/f_launches_1024.txt
for RX in k_edges_tma k_paint_band k_replay "k_path_walk<0>" "k_path_walk<1>" k_expand k_event_setup k_vw_build k_band_ccl "k_run_finish<0>" "k_run_finish<1>" k_dec_chain\\b k_dec_mark k_node_init k_run_init; do
  NAME=$(echo "$RX" | tr -c 'A-Za-z0-9_' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 1 -c 1 -o gpurun_out/ffull_${NAME} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/f_ncu_full_${NAME}.log 2>&1
done
ls -la gpurun_out/ffull_*.ncu-rep | wc -l
tail -c 400 gpurun_out/f_bench1024_n1.json
