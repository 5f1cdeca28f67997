# Multi-GPU evidence (run under `gpurun --gpus N`): the NCCL parity tests (N ranks == 1 rank, bit for bit), bench.py at N ranks
# (configs[2]: 20,000 images strong-scaled, with the secondary block) and the concurrent host->device copy ceiling.
# Usage: bash tools/evidence_multi.sh <N> [prefix]; outputs under gpurun_out/.
N=${1:-2}
P=${2:-r02}
set -x
python -m pytest tests/test_gpu_round2.py -m gpu -q -k nccl 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${P}_bench_n$N.json 2> gpurun_out/${P}_bench_n$N.err; echo "bench N=$N rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/h2d_probe.py 2>/dev/null | tail -1 \
    > gpurun_out/${P}_h2d_probe_n$N.json
