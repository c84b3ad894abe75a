#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash tools/gpu/verify.sh
bash tools/gpu/sanitize.sh
