#!/bin/bash
# all GPU tests + smoke of the current build
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12
