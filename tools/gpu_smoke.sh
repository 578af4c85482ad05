#!/bin/bash
O=gpurun_out/${1:-r01smoke}
mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -14 $O/smoke.log
