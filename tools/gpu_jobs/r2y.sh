timeout 900 python -m pytest tests/test_host_api.py -q -x 2>&1 | grep -v "Warning: Particle" | tail -3
