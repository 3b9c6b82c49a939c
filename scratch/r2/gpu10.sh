#!/bin/bash
cd $GRAFT_REPO_ROOT
echo "== small_first"; timeout 600 python scratch/group_check.py 120000 1 small_first 2>&1 | grep -v "^  File\|^    \|\^\^\|debug\] create\|debug\] aligner" | tail -3
echo "== all"; timeout 600 python scratch/group_check.py 120000 1 all 2>&1 | grep -v "^  File\|^    \|\^\^\|debug\] create\|debug\] aligner" | tail -4
bash scratch/r2/gpu08.sh
