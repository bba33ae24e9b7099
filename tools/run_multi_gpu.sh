#!/bin/bash
# Multi-GPU measurements of one 8xB200 box (run under `gpurun --gpus 8`): the ABR incremental step under DDP (config 2),
# the headline RoI path (config 1), the FPN Pooler (config 5) and the ABR paste (config 4) at 2/4/8 GPUs.
# Results land in gpurun_out/mg_*.json (one JSON line each, after NCCL's version banner if any).
run() {  # run <nproc> <port> <out> <script> [args...]
  local n=$1 port=$2 out=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@" \
      > gpurun_out/$out.json 2> gpurun_out/$out.err
  echo "$out rc=$?"
}
mkdir -p gpurun_out
run 8 29601 mg_abr_ours_n8 tools/abr_step.py --arm ours --steps 8 --warmup 3
run 4 29602 mg_abr_ours_n4 tools/abr_step.py --arm ours --steps 8 --warmup 3
run 8 29603 mg_abr_reference_n8 tools/abr_step.py --arm reference --steps 8 --warmup 3
run 8 29604 mg_bench_n8 bench.py --gpus 8 --steps 20 --warmup 3
run 4 29605 mg_bench_n4 bench.py --gpus 4 --steps 20 --warmup 3
run 8 29606 mg_fpn_n8 bench.py --gpus 8 --workload fpn --steps 50
for n in 2 4 8; do run $n 2961$n mg_paste_n$n bench.py --gpus $n --workload paste --steps 10; done
nvidia-smi topo -m > gpurun_out/mg_topo.txt 2>&1
