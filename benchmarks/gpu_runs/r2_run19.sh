set -x
timeout 600 python -m pytest tests/test_networks.py -m gpu -q --tb=line -rf -s -k "cuda_graph_phases" 2>&1 | grep -E "^FAILED|passed|failed|differs|upfirdn2d launches|unseen" | head -12
