# one GPU: smoke, GPU test-suite, bench at the N = 1 workload and at one GPU's share of the N = 8 workload
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 ${BENCH_FLAGS} > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; grep -v "Warning\|^  \"\"\"" gpurun_out/c1_n1.err | tail -5
python bench.py --steps 20 --warmup 5 --curves 128 --no-cpu-baseline > gpurun_out/c1_c128.json 2> gpurun_out/c1_c128.err; tail -c 300 gpurun_out/c1_c128.err
python tools/bench_summary.py gpurun_out/c1_n1.json gpurun_out/c1_c128.json
