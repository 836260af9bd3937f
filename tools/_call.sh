set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "sinkhorn or row_corr or ffn or mlp2" -s 2>&1 | grep -E "row_corr split|passed|failed|FAILED|Error|mismatch" > gpurun_out/r2/t_new.log; cat gpurun_out/r2/t_new.log
python -m pytest tests/test_parity_gpu.py -q -m gpu -s 2>&1 | grep -E "EPE|flips|golden|passed|failed|FAILED|Error" > gpurun_out/r2/t_parity.log; cat gpurun_out/r2/t_parity.log
python tools/profile_layers.py --batch 8 --top 400 -o gpurun_out/r2/layers_b8_c4.md > /dev/null 2> gpurun_out/r2/layers_b8.err; tail -2 gpurun_out/r2/layers_b8.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_b.json 2> gpurun_out/r2/bench_b.err; cat gpurun_out/r2/bench_b.json | cut -c1-300; tail -3 gpurun_out/r2/bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2/launches_b8_c4.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2/ncu_b.log 2>&1; tail -2 gpurun_out/r2/ncu_b.log
