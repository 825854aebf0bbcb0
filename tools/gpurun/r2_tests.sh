python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -m bayesflare_b200.longseries --samples 130000 2>&1 | tail -1
python -m bayesflare_b200.longseries --samples 1040000 2>&1 | tail -1
