set -x
CMD="python bench.py --steps 1 --warmup 3 --sub-batches 2 --skip c1,c3,c5,e2e,variants --no-cpu-baseline"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo bench rc=$?
$CMD > gpurun_out/r02b_plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:'stft_cc_rr|gcc_fft|srp_gather_ws|shift_stack_vec|topk_kernel|peak_flag|select_kernel' -s 14 -c 10 -f -o gpurun_out/prof_r02b_main $CMD > gpurun_out/r02b_ncu_full.log 2>&1; echo full rc=$?
