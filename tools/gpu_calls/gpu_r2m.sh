#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python tools/timeline.py --csv $O/r2m_timeline.csv > $O/r2m_timeline.txt 2>&1; echo "timeline rc=$?"
head -3 $O/r2m_timeline.txt; tail -2 $O/r2m_timeline.txt
timeout 200 python tools/forward_phases.py 2>&1 | tail -5
