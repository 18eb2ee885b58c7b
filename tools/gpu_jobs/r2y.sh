timeout 900 python -m pytest tests/test_host_api.py tests/test_abi.py -q -x 2>&1 | grep -v "Warning: Particle" | tail -6
