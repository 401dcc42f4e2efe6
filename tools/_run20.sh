python tools/kbench.py | tail -1
python tools/kbench.py 1000000 LANDSAT8-OLI 3 | tail -1
python tools/pcie_probe.py > gpurun_out/r02_pcie_probe.json; cat gpurun_out/r02_pcie_probe.json
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
