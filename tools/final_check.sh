#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 2>&1 | tail -4
timeout 600 python tools/sanitize_small.py 2>&1 | tail -3
