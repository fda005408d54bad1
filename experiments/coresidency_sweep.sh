#!/bin/bash
# Co-residency sweep (round 2): shift-stack variant x STFT CTA cap, quick bench lines into gpurun_out/cores_*.json
mkdir -p gpurun_out
Q="--steps 6 --warmup 3 --skip c1,c3,c5,e2e,variants --no-cpu-baseline"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py $Q > gpurun_out/cores_$name.json 2> gpurun_out/cores_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cores_$name.json").read().strip().splitlines()[-1])
    p=d["stages"]["pipelined"]
    print("$name", "ms/step %.3f"%d["ms_per_step"], "stack_kernel_frac %.3f"%d["roofline"]["frac"], "step_frac %.3f"%d["roofline"]["step"]["frac"],
          "score_busy %.3f stack_busy %.3f wall %.3f"%(p["score_busy_ms_per_sub_batch"],p["stack_busy_ms_per_sub_batch"],p["wall_ms_per_sub_batch"]),
          "serial score %.3f prune %.3f stack %.3f"%(d["stages"]["srp_score_ms"],d["stages"]["prune_ms"],d["stages"]["shift_stack_ms"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/cores_$name.err").read()[-1500:])
PY
}
HI="ASW_STACK=persist ASW_BENCH_STACK_PRIORITY=-1 ASW_BENCH_SCORE_PRIORITY=0"
run persist_hi_c71 $HI ASW_CARVE=71
run persist_hi_c100 $HI ASW_CARVE=100
run persist_hi_c71_stft1 $HI ASW_CARVE=71 ASW_STFT_CTAS=1
run persist_lo_c71 ASW_STACK=persist ASW_CARVE=71
run persist_hi_c50 $HI ASW_CARVE=50
