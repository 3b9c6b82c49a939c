#!/bin/bash
cd $GRAFT_REPO_ROOT
echo "== twice"; MMG_DEBUG_SYNC=1 timeout 600 python scratch/group_check.py 120000 1 twice 2>&1 | grep -v "^  File\|^    \|\^\^\|debug\] create\|debug\] aligner" | tail -6
