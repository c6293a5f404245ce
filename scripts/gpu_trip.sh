#!/bin/bash
# generic GPU trip: $1 = pytest -k filter ('' = all), $2 = engines, $3 = extra command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SRK_TEST_ENGINES=${2:-tcgen05,mma_sync}
if [ -n "$1" ]; then K=(-k "$1"); else K=(); fi
timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests/test_gpu.py -m gpu -q -x "${K[@]}" 2>&1 | tail -120 > gpurun_out/pytest.log
grep -E "passed|failed|^FAILED|^ERROR|Error|assert " gpurun_out/pytest.log | head -60
if [ -n "$3" ]; then eval "$3"; fi
