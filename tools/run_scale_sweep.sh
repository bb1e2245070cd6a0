#!/bin/bash
# BASELINE.json configs 4 and 5 at G GPUs of one box: N = 1M strong, N = 4M strong, weak (N = 1M * sqrt(G)), each with
# bench.py's parity key; then the one-process multi-device handle on config 4.  Usage: tools/run_scale_sweep.sh G
# Writes gpurun_out/r02_scale_g${G}_*.json (copied to profiles/ by the builder).
G=$1
run() { # tag, extra args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G "$@" > gpurun_out/r02_scale_g${G}_${tag}.json 2> gpurun_out/r02_scale_g${G}_${tag}.err
  echo "g=$G $tag rc=$? $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_g${G}_${tag}.json").read().strip().splitlines()[-1])
    print(d["config"]["n_particles"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"], round(d["parity"]["max_dv_over_tol"],3), d["parity"]["ranks_covered"], "frac", round(d["roofline"]["frac"],4))
except Exception as ex:
    print("no line:", ex)
PY
)"
}
run 1m --steps 5 --warmup 3
run 4m --particles 4194304 --steps 3 --warmup 3 --e2e-steps 1 --no-other-cloud
run weak --weak --steps 3 --warmup 3 --e2e-steps 2 --no-other-cloud
timeout 300 python tools/multi_handle_bench.py $G > gpurun_out/r02_multi_handle_g${G}.json 2> gpurun_out/r02_multi_handle_g${G}.err
echo "multi handle g=$G rc=$? $(cut -c1-600 gpurun_out/r02_multi_handle_g${G}.json)"
