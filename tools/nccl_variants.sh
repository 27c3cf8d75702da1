#!/bin/bash
# 2-GPU experiment: where does the N >= 2 step-time penalty come from?  (run under gpurun --gpus 2)
run() { # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus 2 --steps 8 --warmup 3 --parity-check 0 "$@" > gpurun_out/nv_$name.json 2> gpurun_out/nv_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/nv_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"], 1), "utt/s", round(d["ms_per_step"], 2), "ms")
except Exception as e:
    print("$name", "ERR", e)
P
}
run default NCCL_DEBUG=INFO --
run onebucket X=1 -- --bucket-mb 4096
run ch4 NCCL_MAX_NCHANNELS=4 --
run ch8 NCCL_MAX_NCHANNELS=8 --
run eager X=1 -- --graph 0
grep -h "NVLS\|via P2P\|via NVL\|Connected\|channels\|nChannels" gpurun_out/nv_default.err | sort | uniq -c | sort -rn | head -12
