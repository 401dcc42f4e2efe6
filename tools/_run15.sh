bash tools/kbench_all.sh
python tools/lidf_parity_scale.py 1000000 2>/dev/null
python -m pytest tests -m gpu -q -x -k "leafangles or lidf" 2>&1 | tail -3
