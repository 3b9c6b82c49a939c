#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python scratch/group_check.py 60000 0 2>&1 | tail -5
MMG_DEBUG_SYNC=1 timeout 600 python scratch/group_check.py 120000 1 2>&1 | tail -6
timeout 600 python scratch/group_check.py 120000 1 2>&1 | tail -6
