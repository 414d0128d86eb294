#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 70 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_agent.py tests/test_gpu_product_paths.py tests/test_gpu_ppo_c.py tests/test_gpu_returns_variance.py -q -x 2>&1 | tail -3
