set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "sinkhorn" 2>&1 | tail -3
for cs in 4 8; do python tools/bench_dispinit.py --cs $cs 2>&1 | grep -E "Sinkhorn"; done | tee gpurun_out/r2/bench_sk9.log
S2M2_SK_NT=512 python tools/bench_dispinit.py --cs 4 2>&1 | grep -E "Sinkhorn" | tee -a gpurun_out/r2/bench_sk9.log
python tools/bench_dispinit.py 2>&1 | grep -E "Sinkhorn" | tee -a gpurun_out/r2/bench_sk9.log
python tools/bench_dispinit.py --model S --batch 8 2>&1 | grep Sinkhorn | tee -a gpurun_out/r2/bench_sk9.log
