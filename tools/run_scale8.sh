#!/bin/bash
# One 8-GPU box: the default (driver-style) run, the grid A/B for BASELINE configs[2] (strong scaling) and
# BASELINE configs[4] (10M x 128, 256-query batch, N = 8192 vs 16384).  Outputs under gpurun_out/.
# Every run is under `timeout` (RUN_LIMIT seconds, default 420: the default run also records the strong and configs[4] stages): a multi-GPU call is charged N x its wall time, and one
# hung collective in round 2 spent 168 GPU-minutes before anyone could look.
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29600
run() { # name, args...
  name=$1; shift
  P=$((P+1))
  timeout ${RUN_LIMIT:-420} $T --master-port $P bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.log
  echo "== $name rc=$?"; grep -E "ms/step|e2e [0-9]|rror|verif|response share" gpurun_out/$name.log | grep -v "host ms" | tail -n $((3*N+4))
}
run r2_n${N}_default --steps 30 --warmup 3 --no-cpu-baseline
for g in ${GRIDS:-8x1 4x2 2x4 1x8}; do
  run r2_n${N}_c2_grid$g --config sift1m_nlist4096_nprobe64 --grid $g --steps 30 --warmup 3 --no-cpu-baseline --no-parity
done
if [ -z "$SKIP10M" ]; then
run r2_n${N}_10m_n8192_grid2x4 --config synth10m_nlist16384 --grid 2x$((N/2)) --steps 10 --warmup 3 --no-cpu-baseline --no-parity
run r2_n${N}_10m_n8192_grid${N}x1 --config synth10m_nlist16384 --grid ${N}x1 --steps 10 --warmup 3 --no-cpu-baseline --no-parity
run r2_n${N}_10m_n16384_grid2x4 --config synth10m_nlist16384 --poly-degree 16384 --grid 2x$((N/2)) --steps 10 --warmup 3 --no-cpu-baseline --no-parity
fi
df -h /dev/shm | tail -1; free -g | head -2; nproc
