python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4
python -m bayesflare_b200.sweep --realisations 100000 --device-sim 2>/dev/null | tail -1 | cut -c1-330
