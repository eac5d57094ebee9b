set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r4_bench_B512.json 2> gpurun_out/r4_bench_B512.err
python bench.py --workload b48 > gpurun_out/r4_bench_B48_bf16.json 2> gpurun_out/r4_bench_B48_bf16.err
python bench.py --workload b48 --precision fp32 --no-cpu-baseline > gpurun_out/r4_bench_B48_fp32.json 2> gpurun_out/r4_bench_B48_fp32.err
python scripts/bench_b48.py > gpurun_out/r4_bench_b48_graph.json 2> gpurun_out/r4_bench_b48_graph.err
# launch lists (each command first without ncu)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r4_plain512.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r4_launches512.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r4_ncu512.log 2>&1
python bench.py --workload b48 --precision fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r4_plain48f.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r4_launches48f.csv python bench.py --workload b48 --precision fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r4_ncu48f.log 2>&1
# full capture of the score GEMM (6 piece products) of the fp32 step: first acc_gemm launch after the warm-up
ncu --set full --clock-control none --import-source on -k regex:acc_gemm_kernel -s 8 -c 1 -o gpurun_out/r4_f32_gemm python bench.py --workload b48 --precision fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r4_ncu_gemm.log 2>&1
ls -la gpurun_out | tail -20
