set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/policy_pytest.txt
timeout 500 python scripts/sweep.py --policy 2>&1 | tee gpurun_out/policy_matrix.txt
