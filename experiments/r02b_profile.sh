#!/bin/bash
# Round-2 (second pass) evidence bundle, one B200:  gpurun --timeout 1500 -- 'bash experiments/r02b_profile.sh'
# 1. the full bench line; 2. the plain run of the profiled command (must exit 0 without ncu); 3. the ncu launch list of the
# same command; 4. ncu --set full captures of the scoring kernels and of the shift-stack launch (traffic.json).
CMD="python bench.py --steps 1 --warmup 3 --sub-batches 2 --skip c1,c3,c5,e2e,variants --no-cpu-baseline"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo bench rc=$?
$CMD > gpurun_out/r02b_plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:'stft_cc_rr|gcc_fft|srp_gather_ws|topk_kernel|peak_flag|select_kernel' -s 12 -c 6 -f -o gpurun_out/prof_r02b_main $CMD > gpurun_out/r02b_ncu_full.log 2>&1; echo full rc=$?
ncu --set full --clock-control none --import-source on -k regex:shift_stack_vec -s 2 -c 2 -f -o gpurun_out/prof_r02b_stack $CMD > gpurun_out/r02b_ncu_stack.log 2>&1; echo stack rc=$?
