python tools/kbench.py | tail -1
SPART_NO_UNIFORM=1 python tools/kbench.py | tail -1
python tools/kbench.py 1000000 LANDSAT8-OLI 3 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
