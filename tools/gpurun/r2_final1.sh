# final check of the round: smoke(), the whole GPU suite, the bench line
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
( time python bench.py ) > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.log
