#!/bin/bash
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1
head -60 gpurun_out/host_profile.log
