# final check of the round: smoke(), the whole GPU suite, the bench line, the weighted configuration through the API
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
( time python bench.py ) > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err; cut -c1-200 gpurun_out/final_bench.log
python bench_configs.py 6 > gpurun_out/config6_r2.jsonl 2> gpurun_out/config6_r2.err; tail -c 300 gpurun_out/config6_r2.err; cut -c1-900 gpurun_out/config6_r2.jsonl
