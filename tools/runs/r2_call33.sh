#!/bin/bash
# round 2, GPU call 33 (1 GPU): smoke() on the committed tree
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
