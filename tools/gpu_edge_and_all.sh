#!/bin/bash
# edge-case scenes first (own log), then the whole GPU suite and smoke()
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_case" > $O/pytest_edge.log 2>&1; tail -25 $O/pytest_edge.log
( time timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_edge_case_scenes ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
