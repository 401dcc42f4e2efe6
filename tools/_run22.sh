python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/e2e_sweep.py 2>/dev/null | grep -v "^{" | grep "262144"
