#!/bin/bash
# last check of the in-tree product library: smoke() and the sharded world-1 tests (every shape of the pipelined step)
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k world1 2>&1 | tail -2
